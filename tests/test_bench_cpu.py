"""The CPU (reference) arm of bench.py prints ONE JSON line that follows the driver's contract -- runs without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
              'data', 'config', 'e2e', 'cpu_baseline', 'impl'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['steps'] == 1 and d['higher_is_better'] is True
    assert d['unit'] == 'images/sec' and d['metric'].startswith('images/sec ViT-ResNAS-Tiny supernet train step')
    assert d['value'] > 0 and d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
