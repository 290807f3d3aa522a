"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the unmodified reference network — /root/reference or its
staged copy under oracle/_ref/pyref — with custom CUDA disabled, timed on the host cores; the oracle port only where no reference tree is reachable) prints ONE
JSON line with the keys the driver reads, and the CUDA arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(args, timeout=900):
    env = dict(os.environ, OMP_NUM_THREADS='8')
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, cwd=ROOT, env=env, timeout=timeout)


def test_reference_arm_json_line():
    r = _run(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '1', '--ref-frames', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'synthesis_fwd_bwd_frames_per_sec_256' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['n_gpus'] == 1
    from oracle import ref_loader
    assert d['cpu_baseline']['kind'] == ('reference' if ref_loader.available() else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'], capture_output=True,
                       text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_cuda_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--steps', '1', '--warmup', '1', '--no-cpu-baseline'], timeout=300)
    assert r.returncode != 0 and 'no CPU path' in (r.stderr + r.stdout)
