"""The `bench.py --impl reference` arm runs on the host (oracle port of the path, the CPU baseline) -- so the JSON contract of the
bench line can be checked without a GPU: metric / unit / config of BASELINE.json, the e2e and cpu_baseline objects."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith('{')]
    assert len(lines) == 1                                           # ONE JSON line
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0
    assert d['unit'] == 'Msamples/s' and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    for token in ('Msamples/sec', '960', '540', '24spp'):               # BASELINE.json: "rendered Msamples/sec (& Mpix/sec) at 960x540x24spp; ..."
        assert token in d['metric'] and token in base['metric']
    assert d['mpix_per_s'] > 0                                         # the "(& Mpix/sec)" half of the metric
    assert d['value'] > 0 and abs(d['ms_per_step'] * 1e-3 * d['value'] * 1e6 - 64 * 64 * 24) < 1e-3 * 64 * 64 * 24   # value = samples / time
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['sample'] and abs(cb['value'] - d['value']) < 1e-9
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0 and abs(e['value'] - d['value']) < 1e-9 and e['unit'] == d['unit']


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: the driver launches the reference arm like the product arm; rank 0 alone runs it and prints, the others exit 0."""
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29541', os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '1', '--warmup', '0'], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
