#!/usr/bin/env python
"""bench.py -- time-steps/s of the hot path at 1M DOFs (BASELINE.json metric).

Headline workload (config.workload): config 2 of BASELINE.json, lid-driven-cavity/comp_LDC_mesh_convergence.py -- the
compressible Oldroyd-B lid-driven-cavity time step (LPS indicator, 6 assemble+solve sub-steps, the rho1 projection,
2 functionals) on the synthetic refined cavity mesh n=192: 73 728 cells, dim W + dim Zc + dim Q = 999 558 DOFs
(SURVEY.md 8a/8d), reference parameters (Re=10, We=0.25, Ma=0.001, dt=0.001), Krylov tolerances = the reference's
(PETSc default rtol 1e-5), started at t=0.45 where the lid ramp 8(1+tanh(8t-4)) is O(1) so every solve does real work.

  python bench.py [--gpus N --steps K --warmup W]      our arm (CUDA); N>1 under torchrun
  python bench.py --impl reference [...]                reference arm: the CPU restatement of the dolfin/PETSc path
  python bench.py --config cfg1|cfg3|cfg3pre|cfg4|cfg5  the other BASELINE configs at ~1M DOFs as the timed workload

N>1 (torchrun): `value` is the independent-ensemble mode (every rank advances its own member of the
We/Ma sweep, no data-path collective, weak scaling).  The same run then times the MESH-PARTITIONED mode (8e.2, strong
scaling, fused peer-memory Krylov kernels) on the BASELINE config-3 and config-2/5-style large meshes and reports it
under "partition" in the same JSON line, with the 1-GPU time of the same mesh and the parity against it.
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
CONFIGS = ("cfg1", "cfg2", "cfg3", "cfg3pre", "cfg4", "cfg5")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=CONFIGS)
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=N_MESH,
                    help="mesh size (use --mesh-n under torchrun: its parser treats --n as an abbreviation of its own flags)")
    ap.add_argument("--cpu-n", type=int, default=None, help="mesh size of the CPU legs (default: the bench mesh)")
    ap.add_argument("--cpu-solver", default="bicgstab+ilu0", choices=["bicgstab+ilu0", "lu"],
                    help="CPU legs: the reference's own Krylov path (PETSc BiCGStab + ILU(0), rtol 1e-5) or exact LU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config lines under variants.configs")
    ap.add_argument("--no-partition", action="store_true", help="N>1: skip the mesh-partitioned measurements")
    ap.add_argument("--profile-out", default=None, help="write the per-phase timing table here (json)")
    ap.add_argument("--partition", action="store_true",
                    help="under torchrun: ONLY the mesh-partitioned cfg2 run (tools/partition_bench.py)")
    ap.add_argument("--dt", type=float, default=0.001)
    return ap.parse_args()


def workload_config(cfg, n, dofs=None, cells=None):
    """`config` of the JSON line: identical in both arms (the driver compares them)."""
    text = {
        "cfg1": "cfg1 lid-driven-cavity/incomp_LDC.py time step (incompressible Oldroyd-B), n=%d skewed cavity mesh, "
                "Re=0.5 We=0.25 dt=0.005, solver rtol 1e-5, t0=%.2f" % (n, T_START),
        "cfg2": "cfg2 lid-driven-cavity/comp_LDC_mesh_convergence.py time step on the n=%d skewed cavity mesh, "
                "Re=10 We=0.25 Ma=0.001 dt=0.001, solver rtol 1e-5 (reference default), t0=%.2f" % (n, T_START),
        "cfg3": "cfg3 incompressible_benchmarkEWM.py main-pass time step (ramped journal spin), eccentric annulus "
                "%dx%d, script parameters (We=Ma=1e-6, betav=1-eps, A_0=k=B=0), dt=10 hmin^2, t0=%.2f" % (4 * n, n // 4, T_START),
        "cfg3pre": "cfg3 incompressible_benchmarkEWM.py pre-pass time step (un-ramped spin, dt=hmin^2), eccentric "
                   "annulus %dx%d, script parameters" % (4 * n, n // 4),
        "cfg4": "cfg4 fenepmp_jbp.py time step (FENE-P-MP journal bearing), eccentric annulus %dx%d, CSV member "
                "We=0.1 Ma=0.01 lambda_D=0, dt=0.00125, t0=%.2f" % (4 * n, n // 4, T_START),
        "cfg5": "cfg5 compressible_dgp.py time step (buoyancy-driven compressible viscoelastic flow with temperature), "
                "n=%d square mesh, Ra=100 We=0.1 Ma=0.05 dt=0.0005" % n,
    }[cfg]
    out = {"workload": text}
    if dofs is not None:
        out["dofs"] = int(dofs)
    if cells is not None:
        out["cells"] = int(cells)
    return out


# ---- CPU legs: the oracle port of the reference path (dolfin is not installable: see DESIGN.md) -------
def cpu_steps(n, steps, warmup, solver):
    """time `steps` oracle steps of config 2 on an n x n mesh after `warmup` untimed ones.
    Returns dict(sec_per_step, dofs, cells, cores, iterations)."""
    from oracle import configs, fem
    m = fem.skew_mesh(n)
    c = configs.CompLDC(m)
    c.t = T_START
    if solver == "bicgstab+ilu0":
        c.solver = ("bicgstab", "ilu")  # solve(A, x, b, "bicgstab", "default") in a serial PETSc process
    for _ in range(warmup):
        c.step()
    c.last.pop("iterations", None)
    t0 = time.perf_counter()
    for _ in range(steps):
        c.step()
    sec = (time.perf_counter() - t0) / steps
    dofs = c.S["W"].dim + c.S["Zc"].dim + c.S["Q"].dim
    return dict(sec_per_step=sec, dofs=int(dofs), cells=int(len(m.cells)), cores=1,
                iterations=[int(i) for i in c.last.get("iterations", [])][-6:])


def cpu_baseline_dict(res, n, steps, warmup, solver):
    what = ("PETSc-like BiCGStab + ILU(0) at rtol 1e-5 from the previous solution (oracle/krylov_c.c)"
            if solver == "bicgstab+ilu0" else "exact sparse LU (scipy SuperLU)")
    d = {"value": 1.0 / res["sec_per_step"], "unit": UNIT, "cores": res["cores"], "kind": "port",
         "sample": "oracle (numpy/scipy + C restatement of the dolfin/PETSc path) config-2 time step on the n=%d mesh "
                   "(%d DOFs, %d cells): %d timed step(s) after %d warm-up, %.1f s/step, solves: %s; NOT scaled"
                   % (n, res["dofs"], res["cells"], steps, warmup, res["sec_per_step"], what),
         "n": n, "dofs": res["dofs"], "steps": steps, "warmup": warmup, "solver": solver}
    if res["iterations"]:
        d["solver_iterations_last_step"] = res["iterations"]
    return d


def run_reference(args):
    """Reference arm: the CPU restatement of the reference path on the bench configuration itself (n=192 unless
    --cpu-n).  One step costs ~10^2 s of one host core, so the arm runs min(steps, 1) timed step and no warm-up (the
    first-call costs of numpy/scipy are < 1 % of a step) and REPORTS those counts."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n or args.n
    steps, warmup = 1, 0
    t0 = time.perf_counter()
    res = cpu_steps(n, steps, warmup, args.cpu_solver)
    cb = cpu_baseline_dict(res, n, steps, warmup, args.cpu_solver)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * res["sec_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config("cfg2", n, res["dofs"], res["cells"]),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "requested": {"steps": args.steps, "warmup": args.warmup},
            "note": "dolfin 2019.1.0 / PETSc cannot be installed here (no setup.py/pyproject in the reference, no dolfin "
                    "wheel): this arm times the oracle restatement of the same time step on the same mesh, serial like "
                    "the reference scripts; wall %.1f s" % (time.perf_counter() - t0)}
    print(json.dumps(line))


class CpuBaselineJob:
    """the at-config CPU step of `cpu_baseline` in a child process (one host core, BLAS threads pinned to 1), started
    after the GPU measurements so neither disturbs the other (run concurrently it cut the e2e figure by 2.5x)."""

    def __init__(self, n, solver):
        self.n, self.solver = n, solver
        code = ("import json,sys; sys.path.insert(0, %r); import bench; "
                "print('CPUJOB ' + json.dumps(bench.cpu_steps(%d, 1, 0, %r)))" % (ROOT, n, solver))
        self.p = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                  env=dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1",
                                           CUDA_VISIBLE_DEVICES=""))

    def result(self, timeout=900):
        try:
            out, err = self.p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            self.p.kill()
            return {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": "CPU step on the n=%d mesh did not finish within %d s" % (self.n, timeout)}
        for ln in out.splitlines():
            if ln.startswith("CPUJOB "):
                return cpu_baseline_dict(json.loads(ln[7:]), self.n, 1, 0, self.solver)
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "CPU leg failed: " + err[-300:]}


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


def build_driver(cfg, n, device, member=(0.25, 0.001)):
    """the driver of a BASELINE config at ~1M DOFs for mesh parameter n (n=192: 999 558 / 1 003 776 DOFs), with the
    solver settings of the bench: CG + two-level preconditioner for the SPD pressure systems, BiCGStab + Jacobi else."""
    from fenics_b200 import mesh as pmesh
    from fenics_b200.drivers import dgp, ewm, jbp, ldc
    if cfg == "cfg1":
        d = ldc.IncompLDC(pmesh.Skew_Mesh(n, 1, 1), device=device)
        d.t = T_START
        d.pressure_method = "cg"       # the pure-Neumann pressure matrix is symmetric positive SEMI-definite and the
        d.enable_pressure_coarse(256)  # right-hand side consistent: CG + the null-space-safe two-level preconditioner
        return d
    if cfg == "cfg2":
        d = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1), We=member[0], Ma=member[1], device=device)
        d.t = T_START
    elif cfg in ("cfg3", "cfg3pre"):
        d = ewm.EwmJBP(n_theta=4 * n, n_r=n // 4, prepass=(cfg == "cfg3pre"), device=device)  # script parameters
        d.t = T_START if cfg == "cfg3" else 0.0
    elif cfg == "cfg4":
        d = jbp.FenepmpJBP(n_theta=4 * n, n_r=n // 4, device=device)
        d.t = T_START
    else:
        d = dgp.CompDGP(mesh_resolution=n, device=device)
    d.pressure_method = "cg"   # north_star: CG for the (symmetric) pressure system
    d.enable_pressure_coarse(256)  # ... preconditioned by Jacobi + aggregate coarse correction
    return d


def time_config(cfg, n, device, steps, warmup=3):
    """one per-config line of variants.configs: device-resident steps/s, iterations, diagnostics."""
    import torch
    d = build_driver(cfg, n, device)
    d.check = False
    for _ in range(warmup):
        d.step()
    d.sync() if hasattr(d, "sync") else torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = d.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    dofs = sum(d.S[k].dim() for k in ("W", "Zc", "Q"))
    row = dict(workload_config(cfg, n, dofs, d.mesh.num_cells()), steps=steps, warmup=warmup, ms_per_step=round(ms, 3),
               value=round(1e3 / ms, 2), unit=UNIT,
               solver_iterations=dict(zip(d.solve_labels, [i.iterations for i in d.last_info])),
               diagnostics={k: float(v) for k, v in dict(r).items()})
    del d
    torch.cuda.empty_cache()
    return row


def run_ours(args):
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist
    from fenics_b200._lib import lib

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

    cfg = args.config
    drv = build_driver(cfg, args.n, local, SWEEP[rank % len(SWEEP)])
    m = drv.mesh
    coarse = getattr(drv, "pressure_pc", None) == "jacobi_coarse"
    drv.check = False            # solver outcomes fetched once per step, still raising on failure
    drv.pipeline = True          # ... and without stalling the host: step k+1 is enqueued while step k runs
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
    names = [("w0", "w1"), ("p0", "p1"), ("tau0_vec", "tau1_vec"), ("rho0", "rho1"), ("T0", "T1")]
    names = [(a, b) for a, b in names if hasattr(drv, a) and hasattr(drv, b)]
    ins = [getattr(drv, a) for a, _ in names]
    outs = [getattr(drv, b) for _, b in names]
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
        h_in, h_out = h_out, h_in                         # next step's inputs are this step's results (the pinned
                                                          # buffers swap roles: no host-side copy)
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t2.item()) / 1000.0)

    # ---- variant: extrapolated initial guesses (not the headline: the reference starts from x^n) ----
    variant = {}
    if hasattr(drv, "extrapolate_guess"):
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
        variant["extrapolated_initial_guess"] = {
            "value": world * args.steps / (float(t3.item()) / 1000.0), "unit": UNIT,
            "ms_per_step": float(t3.item()) / args.steps,
            "solver_iterations": dict(zip(drv.solve_labels, [i.iterations for i in drv.last_info])),
            "note": "every solve starts from 2 x^n - x^(n-1) instead of x^n (the reference's nonzero_initial_guess); "
                    "same tolerances; optional (`extrapolate_guess`), not used for `value`/`e2e`"}
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
    nb_on = os.environ.get("FX_NB", "1") != "0"

    def kernel_of(lab):
        if lab.startswith("solve:"):
            if lab.endswith(":bicgstab"):
                return "k_bicgstab_nb" if nb_on else "k_bicgstab"
            return "k_pcg_res" if lab.split(":")[2] == "Q" else ("k_cg_nb" if nb_on else "k_cg")
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
    # DRAM bytes the top kernel really moves: per-iteration and per-launch set-up figures of each system from the
    # committed `ncu --set full` capture of THIS build (profiles/r02_top_kernel_traffic.json, made by
    # tools/ncu_traffic.py), times this run's iteration counts -> traffic per average launch and the DRAM fraction
    traffic = dram_frac = traffic_src = None
    try:
        tk = json.load(open(os.path.join(ROOT, "profiles", "r02_top_kernel_traffic.json")))
        if topk == tk["kernel"] and args.n == tk["n"] and cfg == "cfg2":
            tot = 0.0
            sel = [r_ for r_ in rows if kernel_of(r_["phase"]) == topk and "iterations" in r_]
            for r_ in sel:
                sp_ = r_["phase"].split(":")[2]
                # (the P1 density solve of the step converges in 0-1 iterations on a 37 k-row system: not captured, ~0)
                tot += tk["dram_bytes_per_iteration"].get(sp_, 0.0) * r_["iterations"] + tk["dram_bytes_setup"].get(sp_, 0.0)
            traffic = tot / len(sel)
            dram_frac = tot / (top["ms"] * 1e-3) / 1e9 / peak
            traffic_src = tk["source"]
    except Exception:
        pass
    notes = {"k_pcg_res": "shared-memory-resident pipelined CG: grid-barrier latency bound, DRAM traffic ~0",
             "k_cell_matrix": "fp64 atomic scatter: L2 atomic throughput bound"}
    roof = {"bound": "hbm", "kernel": topk, "launches_per_step": top["launches"], "phases": top["phases"],
            "achieved": top["gbs"], "peak": peak, "unit": "GB/s", "frac": top["gbs"] / peak, "traffic": traffic,
            "dram_frac": dram_frac, "traffic_source": traffic_src,
            "peak_source": peak_src, "launch_ms": top["ms"] / top["launches"],
            "alg_bytes_per_launch": top["alg_bytes"] / top["launches"], "iterations": top["iterations"],
            "note": "`achieved`/`frac` = ALGORITHMIC bytes (SURVEY.md 8d: 12 B per entry of DOLFIN's full cell-block "
                    "pattern) over the measured time; the kernel streams a node-block copy (9 / 8.4 B per stored entry, "
                    "explicit zeros and the eliminated DEVSS rows dropped) and serves the 61 MB W matrix from L2 "
                    "(evict_last policy), so `traffic` << algorithmic bytes, `frac` may exceed 1 (it measures work done "
                    "per second on the reference's terms, not DRAM utilisation) and `dram_frac` (real DRAM bytes / time / "
                    "peak) is the hardware utilisation",
            "other_kernels": {k: {"ms_per_step": round(g["ms"], 3), "achieved": round(g["gbs"], 1),
                                  "frac": round(g["gbs"] / peak, 3), "note": notes.get(k, "")}
                              for k, g in groups.items() if k != topk}}
    # north_star's figure of merit: assembly + SpMV (solve) kernels together, algorithmic bytes over their summed time
    allb, allms = sum(g["alg_bytes"] for g in groups.values()), sum(g["ms"] for g in groups.values())
    roof["assembly_plus_solve"] = {"achieved": allb / (allms * 1e-3) / 1e9, "frac": allb / (allms * 1e-3) / 1e9 / peak,
                                   "frac_of_nominal_8000": allb / (allms * 1e-3) / 1e9 / 8000.0, "ms_per_step": allms,
                                   "kernels": sorted(groups)}
    # north_star (1) "mesh colouring, or atomics where colouring loses": the atomic scatter-add against the
    # deterministic atomic-free row gather (FX_ASM_DETERMINISTIC), same forms, same state, CUDA events over 10 launches
    if cfg == "cfg2":
        from fenics_b200 import _abi as abi
        from fenics_b200 import dolfin_api as dapi
        ab = {}
        for sp_, fid, ten in (("W", abi.FX_FORM_C2_A1, drv.AW), ("Zc", abi.FX_FORM_C2_A4, drv.AZ), ("Q", abi.FX_FORM_C2_A5, drv.AQ)):
            n_, nnz_ = pat[sp_]
            ld = {"W": 15, "Zc": 18, "Q": 3}[sp_]
            by = 8 * nnz_ + m.num_cells() * (48 + 8 * ld) + 8 * n_
            row = {"alg_bytes": by}
            for mode in ("atomic", "deterministic"):
                drv.plan.set_assembly_mode(mode)
                f_ = drv.pr.form(fid)
                for _ in range(3):
                    dapi.assemble(f_, tensor=ten)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(10):
                    dapi.assemble(f_, tensor=ten)
                a1.record()
                torch.cuda.synchronize()
                t_ = a0.elapsed_time(a1) / 10
                row[mode] = {"ms": round(t_, 4), "achieved": round(by / (t_ * 1e-3) / 1e9, 1),
                             "frac": round(by / (t_ * 1e-3) / 1e9 / peak, 3)}
            ab[sp_] = row
        drv.plan.set_assembly_mode("atomic")
        roof["assembly_atomic_vs_deterministic"] = ab
    total_ms = sum(r_["ms"] for r_ in rows)
    share = {}
    for r_ in rows:
        k = r_["phase"].split(":")[0]
        share[k] = share.get(k, 0.0) + r_["ms"] / total_ms
    if args.profile_out and rank == 0:
        json.dump({"rows": rows, "share": share, "total_ms": total_ms}, open(args.profile_out, "w"), indent=1)
    E_k = float(r["E_k"])
    member = dict(drv.p)
    del drv, snap
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs at ~1M DOFs (single GPU; BASELINE.md section 5 is filled from these) ----
    if world == 1 and not args.no_configs and cfg == "cfg2":
        variant["configs"] = {}
        for c in CONFIGS:
            if c == "cfg2":
                continue
            try:
                variant["configs"][c] = time_config(c, args.n, local, max(5, args.steps // 2))
            except Exception as exc:  # a config that fails is reported, not hidden
                variant["configs"][c] = {"error": repr(exc)[:300]}

    # ---- N > 1: the mesh-partitioned mode on the same ranks (strong scaling) ----
    partition = None
    if world > 1 and not args.no_partition:
        from tools import partition_suite
        partition = partition_suite.run(rank, world, local)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, args.n, dofs, m.num_cells()),
                "solver": {"W, Zc": "BiCGStab + Jacobi on the node-block layout (k_bicgstab_nb)" if nb_on else "BiCGStab + Jacobi",
                           "pressure": "CG + %s" % ("Jacobi + 256-aggregate coarse correction" if coarse else "Jacobi"),
                           "member": {k: member.get(k) for k in ("Re", "We", "Ma", "dt") if k in member}},
                "parallelism": "ensemble x%d (one independent run per GPU)" % world if world > 1 else "single GPU",
                "l2": "inputs exceed L2: each step streams ~0.6 GB of freshly assembled CSR data "
                      "(W 153 MB, Zc 183 MB matrices re-assembled and re-read every solve)",
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "roofline": roof,
                "phase_share": {k: round(v, 4) for k, v in share.items()},
                "solver_iterations": dict(zip(labels, iters)),
                "variants": variant,
                "E_k": E_k}
        if partition is not None:
            line["partition"] = partition
        if world == 1 and not args.no_cpu_baseline:  # after every GPU measurement: the CPU leg must not share the host with them
            line["cpu_baseline"] = CpuBaselineJob(args.cpu_n or args.n, args.cpu_solver).result()
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
