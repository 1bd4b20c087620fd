"""CPU study: config 1's stress system A4 (LPS diffusion kapp h c1 grad:grad grows like 1/h against the mass term):
Jacobi-BiCGStab against BiCGStab with the additive two-level preconditioner D^-1 + P (P^T A P)^-1 P^T.
usage: python tools/cfg1_stress_twolevel_study.py [n] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import scipy.sparse as sp

from oracle import fem
from oracle.configs import IncompLDC
from tools.cfg1_iter_study import bicgstab_M
from tools.iter_study import rcb_aggregates

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = IncompLDC(fem.skew_mesh(n))
cfg.t = 0.45
t0 = time.time()
for k in range(steps):
    tau_prev = cfg.tau1_vec.vec.copy()
    cfg.step()
print("n=%d %d steps %.0f s; kapp: max %.3g mean %.3g" % (n, steps, time.time() - t0, cfg.kapp.vec.max(), cfg.kapp.vec.mean()), flush=True)
A, b = sp.csr_matrix(cfg.last["A4"]), cfg.last["b4"]
d = A.diagonal()
Zc = cfg.S["Zc"]
comp = Zc.dof_component()
xy = Zc.tabulate_dof_coordinates()
_, it_j = bicgstab_M(A, b, lambda r: r / d, x0=tau_prev)
print("A4 n=%d: %d rows: bicgstab+jacobi %d its" % (n, A.shape[0], it_j), flush=True)
for k in (64, 256, 1024):
    if k * 3 * 8 > A.shape[0]:
        continue
    agg = rcb_aggregates(xy, k) * 3 + comp
    P = sp.csr_matrix((np.ones(len(b)), (np.arange(len(b)), agg)), shape=(len(b), 3 * k))
    Ei = np.linalg.inv((P.T @ A @ P).toarray())
    M = lambda r: r / d + P @ (Ei @ (P.T @ r))  # noqa: E731
    _, it2 = bicgstab_M(A, b, M, x0=tau_prev)
    print("   two-level, %d aggregates (%d patches x 3 components): %d its" % (3 * k, k, it2), flush=True)
