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


def test_two_devices_in_one_process():
    """The library keeps its streams, events, workspaces and kernel attributes per device: a process that calls the fast-mode
    quadblas_qgemm (tensor path: 193 KB of opted-in shared memory, TMA maps, internal streams) and the pipelined host paths on cuda:0
    and then on cuda:1 gets the same bits on both (round 1 kept them in function statics of the first device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    import qblas_b200 as qb
    import qgen
    from qblas_b200 import quad
    rng = np.random.default_rng(4)
    m, n, k = 1536, 1100, 1024                    # >= 64 MB in total: the pipelined all-host path
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    x = quad.random_quads(rng, k); y0 = quad.random_quads(rng, m)
    outs = []
    try:
        for dev in (0, 1, 0):
            torch.cuda.set_device(dev)
            qb.init()
            res = {}
            for mode in (qb.MODE_FAST, qb.MODE_REFERENCE):
                qb.set_mode(mode)
                C = C0.copy()
                qb.quadblas_qgemm("R", "N", "N", m, n, k, 1.5, A, k, B, n, 0.5, C, n)
                if mode == qb.MODE_FAST:
                    assert qb.oz_last_stats()["pairs"] > 0
                y = y0.copy()
                qb.quadblas_qgemv("R", "N", m, k, 1.0, A, k, x, 1, 0.0, y, 1)
                res[mode] = (C, y, qb.quadblas_qdot(k, x, 1, x, 1))
                dC = torch.from_numpy(C0.view(np.int64)).cuda(dev)
                qb.gemm("R", 256, 512, 256, 1.0, torch.from_numpy(A.view(np.int64)).cuda(dev), k, torch.from_numpy(B.view(np.int64)).cuda(dev), n, 0.0, dC, n)
                torch.cuda.synchronize()
            outs.append(res)
    finally:
        torch.cuda.set_device(0)
        qb.set_mode(qb.MODE_REFERENCE)
    for res in outs[1:]:
        for mode in outs[0]:
            assert (res[mode][0] == outs[0][mode][0]).all() and (res[mode][1] == outs[0][mode][1]).all() and res[mode][2] == outs[0][mode][2]
