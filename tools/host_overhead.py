"""Where the step time goes beyond the kernels: host enqueue time per step (pipelined driver, no sync) against the
GPU time of the same steps, and the per-phase CUDA-event table of one synchronous step.
usage: python tools/host_overhead.py [cfg] [n] [steps]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20

d = bench.build_driver(cfg, n, 0)
d.check = False
d.pipeline = True
for _ in range(5):
    d.step()
d.sync()
torch.cuda.synchronize()
out = {"cfg": cfg, "n": n, "steps": steps}
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    host = []
    for _ in range(steps):
        h0 = time.perf_counter()
        d.step()
        host.append(time.perf_counter() - h0)
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    d.sync()
    out["rep%d" % rep] = dict(gpu_ms_per_step=e0.elapsed_time(e1) / steps, host_enqueue_ms_per_step=1e3 * t_enq / steps,
                              host_ms_first5=[round(1e3 * h, 3) for h in host[:5]],
                              host_ms_last5=[round(1e3 * h, 3) for h in host[-5:]])
d.pipeline = False
d.profile = []
d.check = True
d.step()
torch.cuda.synchronize()
ph = [(lab, a.elapsed_time(b)) for lab, a, b in d.profile]
out["phase_sum_ms"] = sum(t for _, t in ph)
out["phases"] = [(lab, round(t, 4)) for lab, t in ph]
print(json.dumps(out, indent=1))
