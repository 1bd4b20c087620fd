"""torchrun check of the mesh-partitioned mode: K steps of a config on a partition vs the single-GPU run.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [n] [K] [bench_n] [cfg2|cfg3|cfg4|cfg5]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import dgp, ewm, jbp, ldc, ldc_dist  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bench_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = sys.argv[4] if len(sys.argv) > 4 else "cfg2"
GENERAL3 = dict(dt=0.002, Re=7.0, We=0.3, Ma=0.05, betav=0.6, al=0.2, Di=0.004, Vh=0.03, Bi=0.2, th=0.7)
CFG = {"cfg2": (ldc.CompLDC, ldc_dist.DistCompLDC, lambda k: pmesh.Skew_Mesh(k, 1, 1), {}),
       "cfg3": (ewm.EwmJBP, ldc_dist.DistEwmJBP, lambda k: pmesh.annulus_mesh(4 * k, max(2, k // 2), x_1=-0.9), GENERAL3),
       "cfg4": (jbp.FenepmpJBP, ldc_dist.DistFenepmpJBP, lambda k: pmesh.annulus_mesh(4 * k, max(2, k // 2)), dict(We=0.5)),
       "cfg5": (dgp.CompDGP, ldc_dist.DistCompDGP, lambda k: dgp.DGP_Mesh(k, 1, 1), {}),
       "cfg6": (dgp.IncompDGP, ldc_dist.DistIncompDGP, lambda k: dgp.DGP_Mesh(k, 1, 1), dict(Ra=5000.0, We=0.2))}
Serial, Dist, make_mesh, params = CFG[cfg]
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

d = Dist(make_mesh(n), device=local, **params)
s = Serial(make_mesh(n), device=local, **params)
for drv in (d, s):
    drv.solve_rtol = 1e-13
    drv.pressure_method = "cg"
    drv.t = 1.4 if cfg in ("cfg5", "cfg6") else 0.45
worst = 0.0
for k in range(K):
    rd, rs = d.step(), s.step()
    for key in rs:
        if key in ("t", "E_e"):  # E_e: cancellation-limited functional, compared through the fields
            continue
        worst = max(worst, abs(rd[key] - rs[key]) / max(abs(rs[key]), 1e-8))
for name, (gid, val) in d.owned_fields().items():
    g = getattr(s, name).vector().get_local()
    worst = max(worst, float(np.max(np.abs(val - g[gid])) / np.max(np.abs(g))))
t = torch.tensor([worst], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = float(t) < 1e-8
out = dict(check="partitioned %s vs single GPU" % cfg, n=n, steps=K, world=world, max_rel_err=float(t), ok=ok,
           owned_cells=d.lp.nc_owned, local_cells=d.mesh.num_cells(),
           iters=[i.iterations for i in d.last_info], iters_serial=[i.iterations for i in s.last_info])
if bench_n:
    del d, s
    db = Dist(make_mesh(bench_n), device=local, **params)
    db.pressure_method = "cg"
    out["pressure_coarse"] = bool(db.enable_pressure_coarse())
    if cfg != "cfg5":
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
    db.profile = []
    db.step()
    torch.cuda.synchronize()
    ph = {}
    for lab, a, b in db.profile:
        ph[lab] = ph.get(lab, 0.0) + a.elapsed_time(b)
    out["phases_ms"] = {k: round(v, 3) for k, v in sorted(ph.items(), key=lambda kv: -kv[1])[:12]}
    out["fused"] = bool(db.ctx.fused)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
