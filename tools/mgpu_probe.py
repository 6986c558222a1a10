"""Probe of the row-sharded qgemm step on N GPUs (run under torchrun): shared residue planes vs element panels, with the tensor
kernel's launch timeline.  usage: torchrun --nproc-per-node N tools/mgpu_probe.py M n k [panel_cols]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import qblas_b200 as qb
from qblas_b200 import dist as qd
from gpu_util import dev_random
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
qb.init(); qb.set_mode(qb.MODE_FAST)
M, n, k = (int(v) for v in sys.argv[1:4]); pw = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
m_loc = M // world
A = dev_random((m_loc * k,), "D113", 100 + rank, dev)
B = dev_random((k * n,), "D113", 7, dev) if rank == 0 else torch.empty((k * n, 2), dtype=torch.int64, device=dev)
buf = qd.SymmetricBuffer(M * n * 16)
C = buf.tensor.view(torch.int64).reshape(M * n, 2)
packed = torch.empty((k * n, 2), dtype=torch.int64, device=dev)
torch.cuda.empty_cache()
for name, shp in (("element panels", None), ("shared planes", qd.SharedPlanes()), ("element panels", None), ("shared planes", qd.SharedPlanes())):
    def step():
        qd.qgemm_row_sharded(M, n, k, 1.0, A, B, 0.0, C, peers=buf, b_panels=pw, b_packed=packed, share_planes=shp)
    step(); step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    per = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        per.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    tl = qb.oz_last_mma_timeline()
    gaps = [tl[i + 1][0] - (tl[i][0] + tl[i][1]) for i in range(len(tl) - 1)]
    mma = sum(d for _, d in tl)
    span = tl[-1][0] + tl[-1][1]
    t = torch.tensor([per[-1][0]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    print(f"[rank {rank}] {name}: step {per[-1][0]:.1f} ms (max over ranks {t.item():.1f}), host {per[-1][1]:.1f} ms, mma sum {mma:.1f} in {len(tl)} launches, span {span:.1f}, "
          f"outside the span {per[-1][0] - span:.1f}, gaps {sum(gaps):.1f} (max {max(gaps) if gaps else 0:.1f}); steps {[round(p[0], 1) for p in per]}", flush=True)
    if shp is not None:
        shp.close()
    dist.barrier()
buf.close()
dist.destroy_process_group()
