'''
N > 1 on real GPUs (skipped on a single-GPU box): scripts/multi_gpu_check.py under torchrun — sharded backup + NCCL all-gather
and sharded run_test must reproduce the single-GPU results bit for bit.  The host-side logic of the same paths is covered on
CPU by tests/test_parallel_cpu.py (gloo, world size 2).
'''
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_match_one_gpu():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29531', os.path.join(ROOT, 'scripts', 'multi_gpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'multi-GPU check OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
