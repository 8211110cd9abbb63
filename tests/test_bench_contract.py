"""bench.py contract checks that need no GPU: the CPU (reference) arm prints one JSON line with the agreed keys, on rank 0
only, and uses the host's cores; the clock sampler degrades gracefully without nvidia-smi."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--net', 'lenet', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = _run({'OMP_NUM_THREADS': '1'})               # what torchrun exports: the arm must still use the host's cores
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
              'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'encrypted_images_per_sec' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['vs_baseline'] is None and 'workload' in d['config'] and d['value'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and 'sample' in cb
    assert cb['cores'] == len(os.sched_getaffinity(0))


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []


def test_clock_sampler_without_nvidia_smi():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    old = os.environ.get('PATH', '')
    os.environ['PATH'] = '/nonexistent'
    try:
        r = s.start().stop()
    finally:
        os.environ['PATH'] = old
    assert set(r) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
