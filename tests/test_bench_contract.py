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


def test_reference_arm_honours_steps_and_prints_the_gpu_arms_workload_string():
    """`same_steps` / `same_config` -- the reference arm runs exactly --steps / --warmup and names the workload
    with the very string the GPU arm prints (bench.workload_string)."""
    import bench
    r = _run(['--impl', 'reference', '--steps', '3', '--warmup', '2', '--cpu-sample', '4'])
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['steps'] == 3 and d['warmup'] == 2
    bench.WORKLOAD = 'mnist_fashion'
    assert d['config']['workload'] == bench.workload_string(1024, 11)


def test_roofline_traffic_comes_from_the_rounds_ncu_summary_or_is_null(tmp_path, monkeypatch):
    import bench
    monkeypatch.setattr(bench, 'ROOT', str(tmp_path))
    assert bench.ncu_traffic('mnist_fashion_bf16_b1024_dgrad') is None
    os.makedirs(tmp_path / 'profiles')
    (tmp_path / 'profiles' / 'ncu_traffic.json').write_text(json.dumps({'k': {'bytes': 123.0, 'source': 'profiles/x.md'}}))
    assert bench.ncu_traffic('k') == 123.0 and bench.ncu_traffic('other') is None
