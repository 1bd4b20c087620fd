#!/usr/bin/env python
"""bench.py -- time-steps/s of the hot path at 1M DOFs (BASELINE.json metric).

Workload (config.workload): config 2 of BASELINE.json, lid-driven-cavity/comp_LDC_mesh_convergence.py --
the compressible Oldroyd-B lid-driven-cavity time step (LPS indicator, 6 assemble+solve sub-steps, the
rho1 projection, 2 functionals) on the synthetic refined cavity mesh n=192: 73 728 cells,
dim W + dim Zc + dim Q = 999 558 DOFs, reference parameters (Re=10, We=0.25, Ma=0.001,
dt=0.001), Krylov tolerances = the reference's (PETSc default rtol 1e-5), started at t=0.45 where the
lid ramp 8(1+tanh(8t-4)) is O(1) so every solve does real work.

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA)
  python bench.py --impl reference [...]                    reference arm: CPU path on the host cores

N>1 (torchrun): independent-ensemble mode -- every rank advances its own member of the
We sweep on its own GPU, no data-path collective, weak scaling; value = total steps/s over all ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "time-steps/s at 1M DOFs (assembly+solve)"
UNIT = "steps/s"
N_MESH = 192
T_START = 0.45
# ensemble members (We, Ma) per rank: values of lid-driven-cavity/flow-parameters.csv rows 2-3 that this mesh / time step
# integrates stably (We = 0.001 at n=192, dt=0.001 blows up in the stress solve: the (1-betav)/We coupling is explicit)
SWEEP = [(0.25, 0.001), (0.1, 0.001), (0.5, 0.001), (1.0, 0.001), (0.25, 0.01), (0.1, 0.01), (0.5, 0.01), (1.0, 0.01)]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=N_MESH,
                    help="mesh size (use --mesh-n under torchrun: its parser treats --n as an abbreviation of its own flags)")
    ap.add_argument("--cpu-n", type=int, default=64, help="mesh size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-phase timing table here (json)")
    ap.add_argument("--partition", action="store_true",
                    help="under torchrun: ONE mesh-partitioned run over all ranks (strong scaling, fused peer-memory Krylov "
                         "kernels) instead of the independent ensemble; same JSON keys, scaling = strong")
    ap.add_argument("--dt", type=float, default=0.001)
    return ap.parse_args()


# ---- CPU arm: the oracle port of the reference path (dolfin is not installable: see DESIGN.md) -------
def cpu_sample(n, steps=1):
    """time `steps` oracle steps of config 2 on an n x n mesh; returns (seconds per step, dofs, cores)."""
    from oracle import configs, fem
    m = fem.skew_mesh(n)
    c = configs.CompLDC(m)
    c.t = T_START
    c.step()  # warm-up (numpy/scipy first-call costs)
    t0 = time.perf_counter()
    for _ in range(steps):
        c.step()
    dt = (time.perf_counter() - t0) / steps
    dofs = c.S["W"].dim + c.S["Zc"].dim + c.S["Q"].dim
    return dt, dofs, 1


def cpu_baseline_dict(n):
    sec, dofs, cores = cpu_sample(n)
    scaled = (1.0 / sec) * dofs / 999558.0
    return {"value": scaled, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle (numpy/scipy port of the dolfin path, exact LU solves) config-2 step on the n=%d mesh "
                      "(%d DOFs): %.2f s/step; value = that rate scaled linearly in DOFs to 999558 "
                      "(optimistic for the CPU: sparse LU is superlinear)" % (n, dofs, sec)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    cb = cpu_baseline_dict(args.cpu_n)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2 comp_LDC_mesh_convergence step, n=192 cavity mesh, 999558 DOFs "
                                   "(measured on a bounded n=%d sample, scaled)" % args.cpu_n},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "dolfin 2019.1.0 cannot be installed here (no setup.py/pyproject in the reference, no dolfin "
                    "wheel); this arm times the oracle port, wall %.1f s" % (time.perf_counter() - t0)}
    print(json.dumps(line))


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        """NVML directly (nvidia_ml_py): ~1 ms per sample instead of a ~100 ms nvidia-smi process, so even a
        0.3 s timed region gets tens of samples.  Row layout = the nvidia-smi query's."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        while not self.stop:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append([str(sm), str(mx)] + ["Active" if why & b else "Not Active" for _, b in bits])
            time.sleep(0.01)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) >= 6:
                    self.rows.append(p)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        if not os.environ.get("FX_BENCH_NOCLOCKS"):
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.th.is_alive():
            self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ---- our arm ----------------------------------------------------------------------------------------------
def spmv_bytes(n, nnz):
    return 12 * nnz + 4 * (n + 1) + 16 * n


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fenics_b200 import mesh as pmesh
    from fenics_b200._lib import lib
    from fenics_b200.drivers import ldc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = pmesh.Skew_Mesh(args.n, 1, 1)
    we_member, ma_member = SWEEP[rank % len(SWEEP)]
    drv = ldc.CompLDC(m, We=we_member, Ma=ma_member, device=local)
    drv.pressure_method = "cg"   # north_star: CG for the (symmetric) pressure system
    coarse = drv.enable_pressure_coarse(256)  # ... preconditioned by Jacobi + aggregate coarse correction
    drv.check = False            # solver outcomes fetched once per step, still raising on failure
    drv.pipeline = True          # ... and without stalling the host: step k+1 is enqueued while step k runs
    drv.t = T_START
    dofs = drv.S["W"].dim() + drv.S["Zc"].dim() + drv.S["Q"].dim()

    for _ in range(args.warmup):
        drv.step()
    drv.sync()
    # both timed regions advance the SAME steps of the same flow: state after the warm-up, restored before region 2
    # (every bound Function: the state fields AND the previous solutions that serve as initial guesses)
    snap = {slot: fn.vector().t.clone() for slot, fn in drv.pr.fields.items()}
    snap_t = drv.t

    def restore():
        drv.sync()
        for slot, v in snap.items():
            drv.pr.fields[slot].vector().t.copy_(v)
        drv.t = snap_t

    # ---- timed region 1: device-resident steps ("value") ----
    barrier()
    l0 = lib.fx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            r = drv.step()
        e1.record()
        barrier()
        drv.sync()               # every step's results are on the host and its solves are checked
    launches = lib.fx_launch_count() - l0
    ms = e0.elapsed_time(e1)
    iters = [i.iterations for i in drv.last_info]
    labels = list(drv.solve_labels)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = world * args.steps / (ms_max / 1000.0)

    # ---- timed region 2: end-to-end through the public API with HOST buffers ("e2e") ----
    restore()
    ins = [drv.w0, drv.p0, drv.tau0_vec, drv.rho0]
    outs = [drv.w1, drv.p1, drv.tau1_vec, drv.rho1]
    h_in = [torch.empty(f.vector().size(), dtype=torch.float64).pin_memory() for f in ins]
    h_out = [torch.empty(f.vector().size(), dtype=torch.float64).pin_memory() for f in outs]
    for h, f in zip(h_in, ins):
        h.copy_(f.vector().t)
    torch.cuda.synchronize()
    h2d = sum(h.numel() * 8 for h in h_in)
    d2h = sum(h.numel() * 8 for h in h_out) + 8 * 8
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        for h, f in zip(h_in, ins):                      # this step's inputs: pinned host -> device
            f.vector().t.copy_(h, non_blocking=True)
        drv.w1.assign(drv.w0)                             # state the previous step left on the device
        drv.tau1_vec.assign(drv.tau0_vec)
        r = drv.step()
        for h, f in zip(h_out, outs):                     # results: device -> pinned host
            h.copy_(f.vector().t, non_blocking=True)
        torch.cuda.synchronize()
        for hi, ho in zip(h_in, h_out):                   # next step's inputs are this step's results
            hi.copy_(ho)
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t2.item()) / 1000.0)

    # ---- variant: extrapolated initial guesses (not the headline: the reference starts from x^n) ----
    restore()
    drv.extrapolate_guess = True
    for _ in range(max(3, args.warmup)):
        drv.step()
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in range(args.steps):
        drv.step()
    e5.record()
    barrier()
    drv.sync()
    t3 = torch.tensor([e4.elapsed_time(e5)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    variant = {"extrapolated_initial_guess": {
        "value": world * args.steps / (float(t3.item()) / 1000.0), "unit": UNIT,
        "ms_per_step": float(t3.item()) / args.steps,
        "solver_iterations": dict(zip(drv.solve_labels, [i.iterations for i in drv.last_info])),
        "note": "every solve starts from 2 x^n - x^(n-1) instead of x^n (the reference's nonzero_initial_guess); "
                "same tolerances; optional (`extrapolate_guess`), not used for `value`/`e2e`"}}
    drv.extrapolate_guess = False

    # ---- per-phase attribution (separate, untimed-for-the-metric pass) + roofline of the top kernel ----
    drv.sync()
    drv.pipeline = False
    drv.profile = []
    drv.check = True
    drv.step()
    torch.cuda.synchronize()
    phases = [(lab, a.elapsed_time(b)) for lab, a, b in drv.profile]
    infos = list(drv.last_info)
    drv.profile = None
    solve_idx = 0
    rows = []
    pat = {s: drv.plan.pattern(s)[:2] for s in ("W", "Zc", "Q")}
    for lab, t_ms in phases:
        row = {"phase": lab, "ms": t_ms}
        if lab.startswith("solve:"):
            sp_ = lab.split(":")[2]
            it = infos[solve_idx].iterations
            solve_idx += 1
            n_, nnz_ = pat[sp_]
            spm = 1 if lab.endswith(":cg") else 2
            vecs = 10 if lab.endswith(":cg") else 14  # vector passes per iteration
            by = (it * (spm * spmv_bytes(n_, nnz_) + vecs * 8 * n_) + spmv_bytes(n_, nnz_) + 6 * 8 * n_)
            row.update(iterations=it, alg_bytes=by, gbs=by / (t_ms * 1e-3) / 1e9)
        elif lab.startswith("assemble_matrix:"):
            sp_ = lab.split(":")[2]
            n_, nnz_ = pat[sp_]
            ld = {"W": 15, "Zc": 18, "Q": 3}[sp_]
            by = 8 * nnz_ + m.num_cells() * (48 + 8 * ld) + 8 * n_  # SURVEY.md 8d (one coefficient vector credited)
            row.update(alg_bytes=by, gbs=by / (t_ms * 1e-3) / 1e9)
        rows.append(row)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # the dominant kernel = the kernel FUNCTION with the largest share of the step; its roofline figure is
    # the algorithmic bytes of all its launches in the step over their CUDA-event time (= average launch)
    def kernel_of(lab):
        if lab.startswith("solve:"):
            if lab.endswith(":bicgstab"):
                return "k_bicgstab"
            return "k_pcg_res" if lab.split(":")[2] == "Q" else "k_cg"
        return "k_cell_matrix" if lab.startswith("assemble_matrix:") else None
    groups = {}
    for r_ in rows:
        k = kernel_of(r_["phase"])
        if k and "alg_bytes" in r_:
            g = groups.setdefault(k, {"ms": 0.0, "alg_bytes": 0, "launches": 0, "iterations": 0, "phases": []})
            g["ms"] += r_["ms"]
            g["alg_bytes"] += r_["alg_bytes"]
            g["launches"] += 1
            g["iterations"] += r_.get("iterations", 0)
            g["phases"].append(r_["phase"])
    for g in groups.values():
        g["gbs"] = g["alg_bytes"] / (g["ms"] * 1e-3) / 1e9
    topk = max(groups, key=lambda k: groups[k]["ms"])
    top = groups[topk]
    traffic = None
    try:  # DRAM bytes of the same kernel from the committed ncu --set full capture, scaled to these launches
        tk = json.load(open(os.path.join(ROOT, "profiles", "r01_top_kernel_traffic.json")))
        if topk == "k_bicgstab" and args.n == 192:  # per average W launch (the capture covers the W systems)
            wl = [r_ for r_ in rows if r_["phase"].startswith("solve:") and ":W:bicgstab" in r_["phase"]]
            traffic = sum(tk["dram_bytes_per_iteration"]["W"] * r_["iterations"] for r_ in wl) / len(wl)
    except Exception:
        pass
    notes = {"k_pcg_res": "shared-memory-resident pipelined CG: grid-barrier latency bound, DRAM traffic ~0",
             "k_cell_matrix": "fp64 atomic scatter: L2 atomic throughput bound"}
    roof = {"bound": "hbm", "kernel": topk, "launches_per_step": top["launches"], "phases": top["phases"],
            "achieved": top["gbs"], "peak": peak, "unit": "GB/s", "frac": top["gbs"] / peak, "traffic": traffic,
            "peak_source": peak_src, "launch_ms": top["ms"] / top["launches"],
            "alg_bytes_per_launch": top["alg_bytes"] / top["launches"], "iterations": top["iterations"],
            "other_kernels": {k: {"ms_per_step": round(g["ms"], 3), "achieved": round(g["gbs"], 1),
                                  "frac": round(g["gbs"] / peak, 3), "note": notes.get(k, "")}
                              for k, g in groups.items() if k != topk}}
    # north_star's figure of merit: assembly + SpMV (solve) kernels together, algorithmic bytes over their summed time
    allb, allms = sum(g["alg_bytes"] for g in groups.values()), sum(g["ms"] for g in groups.values())
    roof["assembly_plus_solve"] = {"achieved": allb / (allms * 1e-3) / 1e9, "frac": allb / (allms * 1e-3) / 1e9 / peak,
                                   "frac_of_nominal_8000": allb / (allms * 1e-3) / 1e9 / 8000.0, "ms_per_step": allms,
                                   "kernels": sorted(groups)}
    total_ms = sum(r_["ms"] for r_ in rows)
    share = {}
    for r_ in rows:
        k = r_["phase"].split(":")[0]
        share[k] = share.get(k, 0.0) + r_["ms"] / total_ms
    if args.profile_out and rank == 0:
        json.dump({"rows": rows, "share": share, "total_ms": total_ms}, open(args.profile_out, "w"), indent=1)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cfg2 lid-driven-cavity/comp_LDC_mesh_convergence.py time step on the n=%d "
                                       "skewed cavity mesh (%d cells, %d DOFs), Re=10 We=%g Ma=%g dt=0.001, "
                                       "solver rtol 1e-5 (reference default), t0=%.2f; pressure solve: CG with %s" %
                                       (args.n, m.num_cells(), dofs, drv.p["We"], drv.p["Ma"], T_START,
                                        "Jacobi + 256-aggregate coarse correction" if coarse else "Jacobi"),
                           "dofs": dofs, "parallelism": "ensemble x%d (one independent run per GPU)" % world
                           if world > 1 else "single GPU",
                           "l2": "inputs exceed L2: each step streams ~0.6 GB of freshly assembled CSR data "
                                 "(W 153 MB, Zc 183 MB matrices re-assembled and re-read every solve)"},
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "roofline": roof,
                "phase_share": {k: round(v, 4) for k, v in share.items()},
                "solver_iterations": dict(zip(labels, iters)),
                "variants": variant,
                "E_k": r["E_k"]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_dict(args.cpu_n)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_partition(args):
    """mesh-partitioned cfg2 step (tools/partition_bench.py is the same measurement as a stand-alone tool)."""
    sys.argv = [sys.argv[0], str(args.n), str(args.steps), str(args.dt)]
    import runpy
    runpy.run_path(os.path.join(ROOT, "tools", "partition_bench.py"), run_name="__main__")


def main():
    args = parse()
    if args.partition:
        return run_partition(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
