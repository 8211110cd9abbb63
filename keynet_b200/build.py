"""Build libkeynet_b200.so in-tree with nvcc for sm_100a (B200).  nvcc cross-compiles without a GPU.
Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header changed) and linked with nvcc -shared."""
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_DIR = os.path.join(_HERE, 'lib')
OBJ_DIR = os.path.join(LIB_DIR, 'obj')
LIB_PATH = os.path.join(LIB_DIR, 'libkeynet_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC,-fvisibility=hidden']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(_HERE, '..', 'include', '*.h'))


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB_PATH, sources() + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    hdrs = _headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdrs)]

    def compile_one(src):
        subprocess.check_call([nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', _obj(src), src])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(compile_one, todo))
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC', '-o', LIB_PATH] + [_obj(s) for s in sources()])
    return LIB_PATH


if __name__ == '__main__':
    print(build(force=True, verbose=True))
