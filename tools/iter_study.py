"""CPU study: Jacobi-BiCGStab iteration counts on the W system of config 2, full system vs the
velocity block alone (the W matrices are block lower-triangular: the DEVSS rows (D - grad u, R) never
feed back into the momentum rows).  Usage: python tools/iter_study.py [n] [steps]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import scipy.sparse as sp
from oracle import fem
from oracle.configs import CompLDC


def bicgstab(A, b, dinv, rtol=1e-5, maxit=10000, x0=None):
    """left-Jacobi BiCGStab, convergence on the preconditioned residual (KSPBCGS default norm)."""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = dinv * (b - A @ x)
    r0 = np.linalg.norm(dinv * b)
    rh = r.copy(); p = r.copy(); rho = rh @ r
    for it in range(1, maxit + 1):
        v = dinv * (A @ p)
        alpha = rho / (rh @ v)
        s = r - alpha * v
        t = dinv * (A @ s)
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


def cg(A, b, dinv, rtol=1e-5, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = dinv * r; p = z.copy(); rz = r @ z
    r0 = np.sqrt((dinv * b) @ (dinv * b))
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        z = dinv * r
        if np.linalg.norm(z) <= rtol * r0:
            return x, it
        rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    from oracle.fem import skew_mesh
    mesh = skew_mesh(n)
    cfg = CompLDC(mesh)
    for _ in range(steps):
        cfg.step()
    W = cfg.S["W"]
    comp = W.dof_component()
    vel = np.where(comp < 2)[0]
    for k in ("1", "2", "7"):
        A, b = sp.csr_matrix(cfg.last["A" + k]), cfg.last["b" + k]
        d = A.diagonal()
        x, it = bicgstab(A, b, 1.0 / d)
        Au = A[vel][:, vel]
        off = abs(A[vel][:, np.where(comp >= 2)[0]]).max()
        xu, itu = bicgstab(Au, b[vel], 1.0 / Au.diagonal())
        asym = abs(Au - Au.T).max() / abs(Au).max()
        print("A%s n=%d: full %d its, velocity block %d its (|A_uD|max=%g, asym=%.2e), |x_u diff|=%.2e"
              % (k, n, it, itu, off, asym, abs(x[vel] - xu).max() / abs(xu).max()))
        if asym < 1e-12:
            print("   cg on velocity block:", cg(Au, b[vel], 1.0 / Au.diagonal())[1])


def rcb_aggregates(xy, k):
    """recursive coordinate bisection of the dof coordinates into k (power of two) aggregates."""
    agg = np.zeros(len(xy), dtype=np.int64)
    parts = [np.arange(len(xy))]
    while len(parts) < k:
        new = []
        for idx in parts:
            c = xy[idx]
            ax = int(np.argmax(c.max(0) - c.min(0)))
            o = idx[np.argsort(c[:, ax], kind="stable")]
            new += [o[:len(o) // 2], o[len(o) // 2:]]
        parts = new
    for g, idx in enumerate(parts):
        agg[idx] = g
    return agg


def deflated_cg(A, b, dinv, agg, rtol=1e-5, maxit=20000):
    """Jacobi-PCG deflated with the piecewise-constant aggregate space W (Saad/Yeung/Erhel/Guyomarc'h)."""
    n, k = len(b), agg.max() + 1
    Wm = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, k))
    AW = (A @ Wm).tocsr()
    E = (Wm.T @ AW).toarray()
    Ei = np.linalg.inv(E)
    x = Wm @ (Ei @ (Wm.T @ b))
    r = b - A @ x
    z = dinv * r
    p = z - Wm @ (Ei @ (AW.T @ z))
    rz = r @ z
    r0 = np.linalg.norm(dinv * b)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        z = dinv * r
        if np.linalg.norm(z) <= rtol * r0:
            return x, it
        rzn = r @ z
        p = z + (rzn / rz) * p - Wm @ (Ei @ (AW.T @ z))
        rz = rzn
    return x, maxit


def pressure_study(n, steps=2):
    from oracle.fem import skew_mesh
    cfg = CompLDC(skew_mesh(n))
    for _ in range(steps):
        cfg.step()
    A, b = sp.csr_matrix(cfg.last["A5"]), cfg.last["b5"]
    dinv = 1.0 / A.diagonal()
    xr, it = cg(A, b, dinv)
    print("A5 n=%d (%d dofs): jacobi-cg %d its" % (n, len(b), it))
    xy = cfg.S["Q"].tabulate_dof_coordinates()
    for k in (16, 64, 256, 1024):
        if k * 8 > len(b):
            continue
        x, itd = deflated_cg(A, b, dinv, rcb_aggregates(xy, k))
        print("   deflated k=%4d: %d its, |x-x_cg|/|x| = %.2e" % (k, itd, np.linalg.norm(x - xr) / np.linalg.norm(xr)))


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "p":
    pressure_study(n, steps)


def block_jacobi_cg(A, b, bs, rtol=1e-5, maxit=20000):
    """PCG with exact solves on diagonal blocks of `bs` consecutive rows."""
    n = len(b)
    A = A.tocsr()
    blocks = [(s, min(n, s + bs)) for s in range(0, n, bs)]
    inv = [np.linalg.inv(A[s:e, s:e].toarray()) for s, e in blocks]

    def M(r):
        z = np.empty_like(r)
        for (s, e), Bi in zip(blocks, inv):
            z[s:e] = Bi @ r[s:e]
        return z
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z
    r0 = np.linalg.norm(M(b))
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        z = M(r)
        if np.linalg.norm(z) <= rtol * r0:
            return x, it
        rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def two_level_cg(A, b, dinv, agg, rtol=1e-5, maxit=20000):
    """PCG with the additive two-level preconditioner M^-1 = D^-1 + P E^-1 P^T (P = aggregate indicators)."""
    n, k = len(b), agg.max() + 1
    P = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, k))
    Ei = np.linalg.inv((P.T @ A @ P).toarray())
    M = lambda r: dinv * r + P @ (Ei @ (P.T @ r))  # noqa: E731
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z
    r0 = np.linalg.norm(M(b))
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        z = M(r)
        if np.linalg.norm(z) <= rtol * r0:
            return x, it
        rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def bicgstab_two_level(A, b, dinv, agg, rtol=1e-5, maxit=10000):
    """left-preconditioned BiCGStab with M^-1 = D^-1 + P (P^T A P)^-1 P^T."""
    n, k = len(b), agg.max() + 1
    P = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, k))
    Ei = np.linalg.inv((P.T @ A @ P).toarray())
    M = lambda r: dinv * r + P @ (Ei @ (P.T @ r))  # noqa: E731
    x = np.zeros_like(b)
    r = M(b - A @ x)
    r0 = np.linalg.norm(M(b))
    rh = r.copy(); p = r.copy(); rho = rh @ r
    for it in range(1, maxit + 1):
        v = M(A @ p)
        alpha = rho / (rh @ v)
        s_ = r - alpha * v
        t = M(A @ s_)
        om = (t @ s_) / (t @ t)
        x += alpha * p + om * s_
        r = s_ - om * t
        if np.linalg.norm(r) <= rtol * r0:
            return x, it
        rho_n = rh @ r
        beta = rho_n / rho * alpha / om
        rho = rho_n
        p = r + beta * (p - om * v)
    return x, maxit
