"""CPU study for config 1 (incomp_LDC.py at We = 0.25, dt = 0.005), the bench's state (t0 = 0.45, a few steps): which
solver / preconditioner choices would cut the iteration counts of its stress (A4) and velocity (A1, A7) systems.
usage: python tools/cfg1_iter_study.py [n] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fem
from oracle.configs import IncompLDC
from tools.iter_study import bicgstab, cg, rcb_aggregates, two_level_cg


def bicgstab_M(A, b, M, rtol=1e-5, maxit=5000, x0=None):
    """left-preconditioned BiCGStab with a general M^-1 (callable)"""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = M(b - A @ x)
    r0 = np.linalg.norm(M(b))
    rh = r.copy(); p = r.copy(); rho = rh @ r
    for it in range(1, maxit + 1):
        v = M(A @ p)
        alpha = rho / (rh @ v)
        s = r - alpha * v
        if np.linalg.norm(s) <= rtol * r0:
            return x + alpha * p, it
        t = M(A @ s)
        om = (t @ s) / (t @ t)
        x += alpha * p + om * s
        r = s - om * t
        if np.linalg.norm(r) <= rtol * r0:
            return x, it
        rho_n = rh @ r
        beta = rho_n / rho * alpha / om
        rho = rho_n
        p = r + beta * (p - om * v)
    return x, maxit


def node_block_jacobi(A, nodes):
    """exact inverses of the diagonal blocks coupling the dofs that share a node; nodes[i] = node id of dof i"""
    order = np.argsort(nodes, kind="stable")
    A = A.tocsr()
    cnt = np.bincount(nodes)
    bs = cnt.max()
    assert (cnt == bs).all()
    idx = order.reshape(-1, bs)
    blocks = np.empty((len(idx), bs, bs))
    for a in range(bs):
        for c in range(bs):
            blocks[:, a, c] = np.asarray(A[idx[:, a], idx[:, c]]).ravel()
    inv = np.linalg.inv(blocks)

    def M(r):
        z = np.empty_like(r)
        z[idx] = np.einsum("nab,nb->na", inv, r[idx])
        return z
    return M


def two_level_cg_masked(A, b, dinv, agg, bc_rows, rtol=1e-5, maxit=20000):
    """PCG with M^-1 = D^-1 + P E^-1 P^T, P = aggregate indicators restricted to the free rows"""
    n, k = len(b), agg.max() + 1
    w = np.ones(n)
    w[bc_rows] = 0.0
    P = sp.csr_matrix((w, (np.arange(n), agg)), shape=(n, k))
    E = (P.T @ A @ P).toarray()
    E[np.diag_indices(k)] += 1e-300
    Ei = np.linalg.pinv(E)
    M = lambda r: dinv * r + P @ (Ei @ (P.T @ r))  # noqa: E731
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z
    r0 = np.linalg.norm(M(b))
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        z = M(r)
        if np.linalg.norm(z) <= rtol * r0:
            return it
        rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return maxit


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    cfg = IncompLDC(fem.skew_mesh(n))
    cfg.t = 0.45
    t0 = time.time()
    for k in range(steps):
        w_prev, tau_prev = cfg.w1.vec.copy(), cfg.tau1_vec.vec.copy()
        w12_prev = cfg.w12.vec.copy()
        cfg.step()
    print("n=%d, %d steps (%.0f s), t=%.3f" % (n, steps, time.time() - t0, cfg.t))
    last = cfg.last
    # ---- stress system -------------------------------------------------------------------------------------------
    A, b = sp.csr_matrix(last["A4"]), last["b4"]
    x0 = tau_prev
    d = A.diagonal()
    S = (A - A.T) * 0.5
    print("A4: %d rows; |skew part|_max / |diag|_min = %.2f" % (A.shape[0], abs(S).max() / abs(d).min()))
    _, it_j = bicgstab_M(A, b, lambda r: r / d, x0=x0)
    Zc = cfg.S["Zc"]
    xy = Zc.tabulate_dof_coordinates()
    _, nodes = np.unique(np.round(xy, 12), axis=0, return_inverse=True)
    Mb = node_block_jacobi(A, nodes.ravel())
    _, it_b = bicgstab_M(A, b, Mb, x0=x0)
    ilu = spla.spilu(A.tocsc(), drop_tol=0.0, fill_factor=1.0)  # not PETSc's ILU(0), but the same flavour
    _, it_i = bicgstab_M(A, b, ilu.solve, x0=x0)
    cnt = [0]

    def cb(_):
        cnt[0] += 1
    Md = spla.LinearOperator(A.shape, lambda r: r / d)
    for m in (30, 60):
        cnt[0] = 0
        xg, info = spla.gmres(A, b, x0=x0, M=Md, restart=m, rtol=1e-5, maxiter=3000 // m, callback=cb, callback_type="pr_norm")
        print("   gmres(%d)+jacobi: %d matvecs (info %d, true rel res %.1e)" % (m, cnt[0], info, np.linalg.norm(b - A @ xg) / np.linalg.norm(b)))
    print("   bicgstab: jacobi %d its (%d matvecs), node-block jacobi %d, spilu %d" % (it_j, 2 * it_j, it_b, it_i))
    it_c = fem.solve_krylov_c(last["A4"], x0.copy(), b, "bicgstab", "ilu", rtol=1e-5)[0]
    print("   C restatement: bicgstab + ILU(0) %d its" % it_c)
    # ---- velocity systems: SPD on the velocity block once the Dirichlet values sit in the initial guess -----------
    W = cfg.S["W"]
    comp = W.dof_component()
    vel = np.where(comp < 2)[0]
    for name, xprev in (("1", w12_prev), ("7", w_prev)):
        A, b = sp.csr_matrix(last["A" + name]), last["b" + name]
        Au = A[vel][:, vel].tocsr()
        bu = b[vel]
        x0 = xprev[vel].copy()
        Az = Au.copy()
        Az.eliminate_zeros()   # bc.apply keeps the pattern: explicit zeros
        bc_rows = np.where((Az.getnnz(axis=1) == 1) & (Az.diagonal() == 1.0))[0]
        x0[bc_rows] = bu[bc_rows]
        free = np.setdiff1d(np.arange(len(bu)), bc_rows)
        Aff = Au[free][:, free]
        asym = abs(Aff - Aff.T).max() / abs(Aff).max()
        du = Au.diagonal()
        _, it_bi = bicgstab_M(Au, bu, lambda r: r / du, x0=xprev[vel])
        # CG from x0 with the boundary values in place: the residual on bc rows is 0 and stays 0
        rhs = bu - Au @ x0
        _, it_cg = cg(Au, rhs, 1.0 / du)
        xyu = W.tabulate_dof_coordinates()[vel]
        res = []
        for k in (64, 256, 1024):
            agg = rcb_aggregates(xyu, k) * 2 + comp[vel]   # one aggregate per (patch, component)
            # bc rows/cols out of the coarse space: give them their own empty treatment by zeroing their P entries
            res.append((k, two_level_cg_masked(Au, rhs, 1.0 / du, agg, bc_rows)))
        print("A%s velocity block: %d rows, asym(free) %.1e: bicgstab+jacobi %d its (%d matvecs); cg+jacobi %d; "
              "cg two-level %s" % (name, len(bu), asym, it_bi, 2 * it_bi, it_cg, res))
