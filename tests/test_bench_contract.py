"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the
driver reads, runs on rank 0 only, and times the oracle's replay of the reference loop body on a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--cpu-frames', '12', '--gpus', '1'], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'frames/sec detect+link at 1228x922' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1 and d['vs_baseline'] is None
    assert d['value'] > 0 and 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value'] and '12 frames' in cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []
