"""Mesh-partitioned cfg2 time step under torchrun: steps/s with CUDA events (max over ranks), iteration
counts and the per-phase split.  world 1 runs the single-GPU driver on the same mesh (the comparison row).
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/partition_bench.py n [steps] [dt] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc, ldc_dist  # noqa: E402

n = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.001
out_path = sys.argv[4] if len(sys.argv) > 4 else None
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh = pmesh.Skew_Mesh(n, 1, 1)
drv = ldc_dist.DistCompLDC(mesh, device=local, dt=dt) if world > 1 else ldc.CompLDC(mesh, device=local, dt=dt)
drv.t = 0.45
drv.pressure_method = "cg"
coarse = bool(drv.enable_pressure_coarse())
for _ in range(3):
    drv.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    r = drv.step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
iters = dict(zip(drv.solve_labels, [i.iterations for i in drv.last_info]))
drv.profile = []
drv.step()
torch.cuda.synchronize()
ph = {}
for lab, a, b in drv.profile:
    key = lab.split(":")[0] if not lab.startswith("solve") else lab
    ph[key] = ph.get(key, 0.0) + a.elapsed_time(b)
if world > 1:
    S = drv.global_spaces
    dofs = sum(S[k].dim() for k in ("W", "Zc", "Q"))
    local_dofs = sum(drv.S[k].dim() for k in ("W", "Zc", "Q"))
else:
    dofs = local_dofs = sum(drv.S[k].dim() for k in ("W", "Zc", "Q"))
row = dict(metric="time-steps/s (assembly+solve), one mesh-partitioned run", value=round(1e3 / float(ms), 2), unit="steps/s",
           scaling="strong", dtype="f64", data="synthetic", steps=steps, warmup=3,
           workload="cfg2 comp_LDC_mesh_convergence.py step, skewed cavity n=%d, dt=%g, rtol 1e-5" % (n, dt), n_gpus=world,
           mode=("mesh partition (RCB), fused peer-memory Krylov kernels, replicated P1 solves" if world > 1 else "single GPU"),
           dofs=dofs, dofs_per_gpu_incl_ghosts=local_dofs, ms_per_step=round(float(ms), 3), steps_per_s=round(1e3 / float(ms), 2),
           dof_steps_per_s=round(dofs * 1e3 / float(ms)), pressure_coarse=coarse, iterations=iters, E_k=float(r["E_k"]),
           phases_ms={k: round(v, 3) for k, v in sorted(ph.items(), key=lambda kv: -kv[1])})
if rank == 0:
    print(json.dumps(row), flush=True)
    if out_path:
        with open(out_path, "a") as f:
            f.write(json.dumps(row) + "\n")
if world > 1:
    dist.destroy_process_group()
