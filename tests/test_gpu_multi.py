"""Multi-GPU parity on real GPUs: needs >= 2 devices (gpurun --gpus 2); skipped on the
1-GPU box.  The host-side partition / collective logic is also covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_hot_path_on_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):   # keep the workers' own report (the assertion message below is truncated by pytest -q | tail)
        with open(os.path.join(out_dir, f"mgpu_worker_w{world}.log"), "w") as f:
            f.write(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-20000:])
    assert r.returncode == 0 and "all passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
