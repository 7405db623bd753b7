"""The loss-scalar exchange on real hardware: needs two GPUs (skipped on a one-GPU box; the driver's round-end `-m gpu`
run is such a box -- `profiles/r02_multi_gpu_check_n2.json` is this test's report from a 2 x B200 box).  Runs
tools/multi_gpu_check.py under torchrun, one rank per GPU: NVLink peer mailboxes and NCCL give the float64-accumulated sum
bit-identically on every rank over 2 000 back-to-back calls and as CUDA-graph replays, and ONE 256-image RetinaNet-COCO
batch sharded with `shard_batch` all-reduces (inside the loss kernel, as a separate launch, through NCCL) to the total a
single GPU computes for the whole batch."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_batch_allreduces_to_the_single_gpu_total():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    env = dict(os.environ, DH_CHECK_BATCH="64")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    report = json.loads([ln for ln in out.stdout.split("\n") if ln.startswith("{")][-1])
    assert report["ok"] and report["status_bits"] == 0
    assert report["comm"]["peer"] or report["comm"]["nccl"]
    for name, chk in report["checks"].items():
        assert chk["ok"], (name, chk)
