"""The gradient all-reduce overlapped with backward (trainval.accum_gradient(last=True) -> parallel.GradBuckets, both
collectives captured in the micro-step's CUDA graph) gives the buffer the single all-reduce of apply_gradient() gives.
Runs tests/overlap_worker.py in a fresh process on a 1-rank NCCL group (the N-rank run of the same worker under torchrun:
profiles/README.md).  /root/reference/dgcnn/trainval.py:59-80."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_overlapped_two_bucket_allreduce_in_graph(cuda):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "overlap_worker.py")], capture_output=True, text=True,
                         timeout=600, cwd=root, env=env)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    assert "OVERLAP_OK world=1" in out.stdout
