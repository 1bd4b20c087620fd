"""Where a mesh-partitioned step spends its time (torchrun, N ranks): ms per step with the solver outcomes fetched
after every solve (check=True) and once per step (check=False), and rank 0's per-phase CUDA-event table of one step
in each mode.  usage: torchrun ... tools/partition_phases.py [cfg2|cfg3] [n] [steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import partition_suite as ps  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d, mesh = ps._mk(case, world, local, n)
out = {"case": case, "n": n, "world": world, "dofs": ps._dofs(d, world)}
for mode in ("check_every_solve", "check_once_per_step"):
    d.check = mode == "check_every_solve"
    ms, _ = ps._timed(d, 3, steps, world)
    d.profile = []
    d.step()
    torch.cuda.synchronize()
    ph = [(lab, a.elapsed_time(b)) for lab, a, b in d.profile]
    d.profile = None
    agg = {}
    for lab, t in ph:
        k = lab.split(":")[0] + (":" + lab.split(":")[2] if lab.startswith("solve:") else "")
        agg[k] = agg.get(k, 0.0) + t
    out[mode] = dict(ms_per_step=round(ms, 3), phase_sum_ms=round(sum(t for _, t in ph), 3),
                     by_kind={k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])},
                     solves={lab: round(t, 3) for lab, t in ph if lab.startswith("solve:")},
                     iterations=dict(zip(d.solve_labels, [i.iterations for i in d.last_info])))
    dist.barrier()
if rank == 0:
    print(json.dumps(out, indent=1))
d.ctx.close()
dist.destroy_process_group()
