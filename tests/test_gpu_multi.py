"""Multi-GPU parity under pytest: N ranks (one process per GPU, torchrun) against the single-rank run of the same library,
for every exchange mode -- halo-sum inside the apply kernel overlapped with the interior elements (default), last-CTA tail,
separate LL kernel, NCCL.  Skipped when the box has fewer than 2 GPUs (the round-end GPU test box has one; the bench line
carries the same checks in its `parity` object for N > 1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("args", [["--p2p-fuse", "2"], ["--p2p-fuse", "1"], ["--p2p-fuse", "0"], ["--comm", "nccl"],
                                  ["--p2p-fuse", "2", "--mesh", "cylinder", "--order", "3"], ["--p2p-fuse", "2", "--order", "6"]])
def test_two_ranks_match_one_rank(args):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + abs(hash(tuple(args))) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI-GPU PARITY: OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
