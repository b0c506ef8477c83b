"""bench.py contract on CPU: the reference arm prints exactly one JSON line with the contract's keys (rank 0) and nothing on the
other ranks; the GPU arm refuses to run without a CUDA device (no CPU fallback of the product path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, env=e, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-sample', '8'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['vs_baseline'] is None and d['value'] > 0 and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--steps', '1'])
    assert r.returncode != 0 and 'no CPU fallback' in r.stderr
