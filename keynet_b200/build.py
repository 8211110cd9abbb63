"""Build libkeynet_b200.so in-tree with nvcc for sm_100a (B200).  nvcc cross-compiles without a GPU."""
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_DIR = os.path.join(_HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libkeynet_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC,-fvisibility=hidden', '-shared']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(_HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + sources()
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force=True, verbose=True))
