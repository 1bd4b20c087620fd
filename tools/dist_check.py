"""torchrun check of the mesh-partitioned mode: K steps of config 2 on a partition vs the single-GPU run.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [n] [K] [bench_n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402
from fenics_b200.drivers.ldc_dist import DistCompLDC  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bench_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

m = pmesh.Skew_Mesh(n, 1, 1)
d = DistCompLDC(m, device=local)
s = ldc.CompLDC(m, device=local)
for drv in (d, s):
    drv.solve_rtol = 1e-13
    drv.pressure_method = "cg"
    drv.t = 0.45
worst = 0.0
for k in range(K):
    rd, rs = d.step(), s.step()
    for key in ("E_k", "E_e"):
        worst = max(worst, abs(rd[key] - rs[key]) / max(abs(rs[key]), 1e-300))
ref = dict(w1=s.w1, p1=s.p1, tau1=s.tau1_vec, rho1=s.rho1)
for name, (gid, val) in d.owned_fields().items():
    g = ref[name].vector().get_local()
    worst = max(worst, float(np.max(np.abs(val - g[gid])) / np.max(np.abs(g))))
t = torch.tensor([worst], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = float(t) < 1e-8
out = dict(check="partitioned cfg2 vs single GPU", n=n, steps=K, world=world, max_rel_err=float(t), ok=ok,
           owned_cells=d.lp.nc_owned, local_cells=d.mesh.num_cells(),
           iters=[i.iterations for i in d.last_info], iters_serial=[i.iterations for i in s.last_info])
if bench_n:
    del d, s
    mb = pmesh.Skew_Mesh(bench_n, 1, 1)
    db = DistCompLDC(mb, device=local)
    db.pressure_method = "cg"
    db.t = 0.45
    for _ in range(2):
        db.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    nb = 5
    for _ in range(nb):
        db.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    out.update(bench_n=bench_n, ms_per_step=(time.perf_counter() - t0) / nb * 1e3,
               bench_iters=dict(zip(db.solve_labels, [i.iterations for i in db.last_info])))
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
