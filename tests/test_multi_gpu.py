"""N > 1 on real GPUs (skipped on single-GPU boxes; the host-side logic is covered by tests/test_distributed_cpu.py with gloo):
tools/check_multi_gpu.py under torchrun -- all-gather upload bit-identical to a host upload, sharded requests identical to
unsharded ones (winner, consensus set, compute()), refit within the summation-order tolerance."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_match_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29571", os.path.join(ROOT, "tools", "check_multi_gpu.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI-GPU CHECK OK" in out.stdout, out.stdout[-3000:]
