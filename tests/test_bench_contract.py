"""CPU checks of bench.py's contract (no GPU): the reference arm prints ONE JSON line with the driver's keys, and the
CUDA arm refuses to run (loudly) without a device -- there is no CPU fallback to time by accident."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(['--impl', 'reference', '--workload', 'mlp', '--steps', '1', '--warmup', '1'],
             env={'KFACB_BENCH_CPU_THREADS': '4'})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['unit'] == 'ms' and line['higher_is_better'] is False
    assert line['value'] > 0 and line['e2e']['value'] == line['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] == 4


def test_reference_arm_on_other_ranks_exits_quietly():
    r = _run(['--impl', 'reference', '--workload', 'mlp', '--steps', '1', '--warmup', '1'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a box without a GPU')
def test_cuda_arm_refuses_to_run_without_a_device():
    r = _run(['--steps', '1', '--warmup', '1'])
    assert r.returncode != 0
    assert 'no CUDA device' in (r.stderr + r.stdout)
