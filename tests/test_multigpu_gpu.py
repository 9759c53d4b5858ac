"""2-GPU hardware test of the data-parallel path (SURVEY section 4 'distributed' row): see tests/dist_worker.py.
Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("enc", ["densenet121_bts", "resnext50_bts"])
def test_ddp_gradients_are_shard_means_and_buffers_follow_rank0(enc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker.py"), enc]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT,
                       env=dict(os.environ, BTS_B200_PRETRAINED="0"))
    if r.returncode != 0 or "DIST_OK" not in r.stdout:
        sys.stderr.write("---- worker stdout ----\n" + r.stdout[-6000:] + "\n---- worker stderr ----\n" + r.stderr[-12000:] + "\n")
    assert r.returncode == 0 and "DIST_OK" in r.stdout, "worker failed (see captured stderr)"
