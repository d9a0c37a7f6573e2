"""Data-parallel gradient exchange on real GPUs (needs >= 2; skipped on a 1-GPU box): scripts/ddp_check.py under torchrun.
The CPU-side logic of the same path (autograd node + reducer, gloo, world size 2) is covered in tests/test_host_logic.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_gradients_identical_on_all_ranks_graph_eager_accumulate():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'scripts', 'ddp_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and 'DDP CHECK OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
