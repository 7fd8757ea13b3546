"""bench.py contract, CPU side: the reference arm (`--impl reference`) prints ONE JSON line with the agreed keys.  Run on a
tiny complex so that the oracle step takes a second."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, DDB200_CPU_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--n-res', '40', '--n-atoms', '8', '--poses', '2'], capture_output=True, text=True, env=env,
                         cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'poses/sec at 20 diffusion steps' and d['unit'] == 'poses/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['dtype'] == 'f32' and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] == 4 and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: rank 0 alone runs and prints the reference line, the other ranks exit 0 without work."""
    env = dict(os.environ, DDB200_CPU_THREADS='2')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29547', os.path.join(ROOT, 'bench.py'),
                          '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1', '--n-res', '40',
                          '--n-atoms', '8', '--poses', '2'], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2
