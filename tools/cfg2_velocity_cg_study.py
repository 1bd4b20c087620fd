"""CPU study: the velocity blocks of config 2's W systems (A1, A2, A7) are SPD on the free dofs (all-Dirichlet
cavity): Jacobi-BiCGStab (what the loop runs) against CG with Jacobi and with the two-level preconditioner.
usage: python tools/cfg2_velocity_cg_study.py [n] [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import scipy.sparse as sp

from oracle import fem
from oracle.configs import CompLDC
from tools.cfg1_iter_study import bicgstab_M, two_level_cg_masked
from tools.iter_study import cg, rcb_aggregates

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = CompLDC(fem.skew_mesh(n))
cfg.t = 0.45
prev = {}
for k in range(steps):
    prev = {"1": cfg.w12.vec.copy(), "2": cfg.ws.vec.copy(), "7": cfg.w1.vec.copy()}
    cfg.step()
W = cfg.S["W"]
comp = W.dof_component()
vel = np.where(comp < 2)[0]
xyu = W.tabulate_dof_coordinates()[vel]
for name in ("1", "2", "7"):
    A, b = sp.csr_matrix(cfg.last["A" + name]), cfg.last["b" + name]
    Au, bu = A[vel][:, vel].tocsr(), b[vel]
    Az = Au.copy()
    Az.eliminate_zeros()
    bc_rows = np.where((Az.getnnz(axis=1) == 1) & (Az.diagonal() == 1.0))[0]
    free = np.setdiff1d(np.arange(len(bu)), bc_rows)
    Aff = Au[free][:, free]
    asym = abs(Aff - Aff.T).max() / abs(Aff).max()
    du = Au.diagonal()
    x0 = prev[name][vel].copy()
    _, it_bi = bicgstab_M(Au, bu, lambda r: r / du, x0=x0)
    x0[bc_rows] = bu[bc_rows]
    rhs = bu - Au @ x0
    # the stopping rule is relative to ||M^-1 b||, not to the initial residual: scale the tolerance accordingly
    scale = np.linalg.norm(bu / du) / max(np.linalg.norm(rhs / du), 1e-300)
    _, it_cg = cg(Au, rhs, 1.0 / du, rtol=1e-5 * scale)
    res = []
    for k in (64, 256, 1024):
        if k * 16 > len(bu):
            continue
        agg = rcb_aggregates(xyu, k) * 2 + comp[vel]
        res.append((2 * k, two_level_cg_masked(Au, rhs, 1.0 / du, agg, bc_rows, rtol=1e-5 * scale)))
    print("A%s n=%d: %d rows, asym(free) %.1e: bicgstab+jacobi %d its (%d matvecs); cg+jacobi %d; cg two-level (k, its) %s"
          % (name, n, len(bu), asym, it_bi, 2 * it_bi, it_cg, res), flush=True)
