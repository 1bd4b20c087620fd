"""CPU: the oracle's C restatement of the reference's linear solver (oracle/krylov_c.c: PETSc KSPBCGS / KSPCG + PCILU(0),
what `solve(A, x, b, "bicgstab", "default")` runs, lid-driven-cavity/incomp_LDC.py:331) pinned against independent
definitions, and the schedule of the GPU kernels (fenics_b200/csrc/fx_ilu.cu) modelled statement by statement against it.

* where ILU(0) has no fill to drop (dense and tridiagonal patterns) it IS the LU factorisation: factors equal a textbook
  Doolittle elimination, and the preconditioned Krylov solve ends in one iteration at the exact solution;
* on a finite-element matrix the returned iterate satisfies PETSc's stopping rule, recomputed here from the factors:
  ||M^-1 (b - A x)|| <= rtol ||M^-1 b||, and agrees with a direct solve;
* the device kernels' index logic (k_ilu_factor's binary search on the row, k_ilu_apply's ticket order and 32-wide ordered
  subtraction) and the stream-ordered BiCGStab / CG state machine (k_ilu_vec / k_ilu_scal: accumulators, half-step exit,
  maxit, iteration stamps), transcribed to Python, reproduce the oracle bit for bit / iterate for iterate.  (The GPU test
  tests/test_gpu_ilu.py checks the kernels themselves.)"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fem


def _klib():
    fem.solve_krylov_c(sp.eye(2, format="csr"), np.zeros(2), np.ones(2))
    return fem._KLIB


def _ilu0(A):
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data, dtype=np.float64)
    lu, dptr = np.zeros_like(va), np.zeros(n, dtype=np.int32)
    rc = _klib().fxo_ilu0(C.c_int(n), *(C.c_void_p(a.ctypes.data) for a in (rp, ci, va, lu, dptr)))
    return rc, rp, ci, va, lu, dptr


def _apply(rp, ci, lu, dptr, r):
    n = len(dptr)
    z = np.zeros(n)
    for i in range(n):
        s = r[i]
        for j in range(rp[i], dptr[i]):
            s -= lu[j] * z[ci[j]]
        z[i] = s
    for i in range(n - 1, -1, -1):
        s = z[i]
        for j in range(dptr[i] + 1, rp[i + 1]):
            s -= lu[j] * z[ci[j]]
        z[i] = s / lu[dptr[i]]
    return z


def _fe_like(n, density, seed, shift=3.0):
    A = sp.random(n, n, density, random_state=seed, format="csr") + sp.eye(n) * shift
    A = (A + 0.3 * A.T).tocsr()
    A.sort_indices()
    return A


def test_ilu0_is_lu_where_nothing_is_dropped():
    rng = np.random.default_rng(0)
    n = 40
    D = rng.standard_normal((n, n)) + n * np.eye(n)
    T = sp.diags([rng.standard_normal(n - 1), 4 + rng.random(n), rng.standard_normal(n - 1)], [-1, 0, 1]).toarray()
    for M in (D, T):
        pat = sp.csr_matrix(M) if M is T else sp.csr_matrix((M.ravel(), np.tile(np.arange(n), n), np.arange(n + 1) * n), (n, n))
        rc, rp, ci, va, lu, dptr = _ilu0(pat)
        assert rc == 0
        LU = M.copy()                                   # Doolittle, no pivoting
        for k in range(n - 1):
            LU[k + 1:, k] /= LU[k, k]
            LU[k + 1:, k + 1:] -= np.outer(LU[k + 1:, k], LU[k, k + 1:])
        got = sp.csr_matrix((lu, ci, rp), (n, n)).toarray()
        mask = sp.csr_matrix((np.ones_like(lu), ci, rp), (n, n)).toarray() > 0
        assert np.max(np.abs(got - LU)[mask]) < 1e-12 * np.abs(LU).max()
        b = rng.standard_normal(n)
        x = np.zeros(n)
        it, rn, bn = fem.solve_krylov_c(pat, x, b, "bicgstab", "ilu", rtol=1e-10)
        assert it <= 1 and np.allclose(x, np.linalg.solve(M, b), rtol=1e-10, atol=1e-12)
    # zero pivot: reported (status -2), like PETSc's "zero pivot in LU factorization"
    Z = sp.csr_matrix(np.array([[0.0, 1.0], [1.0, 1.0]]))
    assert _ilu0(Z)[0] == 1
    with pytest.raises(RuntimeError):
        fem.solve_krylov_c(Z, np.zeros(2), np.ones(2), "bicgstab", "ilu")


@pytest.mark.parametrize("method", ["bicgstab", "cg"])
def test_stopping_rule_and_direct_solution(method):
    A = _fe_like(300, 0.03, 3)
    if method == "cg":
        A = (A @ A.T + sp.eye(300)).tocsr()
        A.sort_indices()
    rng = np.random.default_rng(1)
    b, x0 = rng.standard_normal(300), 0.1 * rng.standard_normal(300)
    rc, rp, ci, va, lu, dptr = _ilu0(A)
    assert rc == 0
    ref = spla.splu(A.tocsc()).solve(b)
    its = {}
    for pc in ("ilu", "jacobi"):
        for rtol in (1e-5, 1e-10):
            x = x0.copy()
            it, rn, bn = fem.solve_krylov_c(A, x, b, method, pc, rtol=rtol)
            its[pc, rtol] = it
            if pc == "ilu":
                assert abs(bn - np.linalg.norm(_apply(rp, ci, lu, dptr, b))) <= 1e-13 * bn
                true = np.linalg.norm(_apply(rp, ci, lu, dptr, b - A @ x))
                assert true <= 1.01 * rtol * bn + 1e-14 * bn and abs(true - rn) <= 1e-3 * rtol * bn + 1e-13 * bn
            assert np.linalg.norm(x - ref) <= 100 * rtol * np.linalg.norm(ref)
    assert its["ilu", 1e-10] < its["jacobi", 1e-10]
    x = ref.copy()                                       # warm start at the solution: nothing to do
    assert fem.solve_krylov_c(A, x, b, method, "ilu", rtol=1e-8)[0] == 0


def test_model_of_the_device_factorisation_and_triangular_solves_is_bit_exact():
    """k_ilu_diag / k_ilu_factor / k_ilu_apply (fx_ilu.cu) with the warps run one after the other in ticket order: same
    index arithmetic, rounded multiply-then-subtract, 32-wide chunks subtracted in lane order."""
    A = _fe_like(200, 0.06, 2)
    rc, rowptr, col, val, luo, dptr = _ilu0(A)
    n = A.shape[0]
    dp = np.full(n, -1)
    for i in range(n):
        for j in range(rowptr[i], rowptr[i + 1]):
            if col[j] == i:
                dp[i] = j
                break
    assert np.array_equal(dp, dptr)
    lu = val.copy()
    for i in range(n):
        r1, di = rowptr[i + 1], dp[i]
        for j in range(rowptr[i], di):
            k = col[j]
            dk = dp[k]
            f = lu[j] / lu[dk]
            lu[j] = f
            if f != 0.0:
                for lane in range(32):
                    for q in range(dk + 1 + lane, rowptr[k + 1], 32):
                        c, lo, hi, p = col[q], j + 1, r1 - 1, -1
                        while lo <= hi:
                            mid = (lo + hi) >> 1
                            if col[mid] == c:
                                p = mid
                                break
                            if col[mid] < c:
                                lo = mid + 1
                            else:
                                hi = mid - 1
                        if p >= 0:
                            lu[p] = lu[p] - f * lu[q]
    assert np.array_equal(lu, luo)
    r = np.random.default_rng(1).standard_normal(n)
    y, z = np.zeros(n), np.zeros(n)
    for t in range(2 * n):
        fwd = t < n
        i = t if fwd else 2 * n - 1 - t
        di = dp[i]
        j0, j1 = (rowptr[i], di) if fwd else (di + 1, rowptr[i + 1])
        src = y if fwd else z
        s = r[i] if fwd else y[i]
        for jb in range(j0, j1, 32):
            terms = [lu[jb + l] * src[col[jb + l]] if jb + l < j1 else 0.0 for l in range(32)]
            for l in range(min(32, j1 - jb)):
                s = s - terms[l]
        if fwd:
            y[i] = s
        else:
            z[i] = s / lu[di]
    assert np.array_equal(z, _apply(rowptr, col, luo, dptr, r))


class _S:
    pass


def _device_schedule(A, prec, method, b, x, rtol, atol, maxit, chunk=4):
    """the launch sequence of fx::la_krylov_ilu0 with k_ilu_vec / k_ilu_scal / k_ilu_spmv / k_ilu_apply as numpy"""
    n = len(b)
    s = _S()
    s.bn = s.tol = s.rho = s.alpha = s.omega = s.beta = s.rn = 0.0
    s.acc, s.done, s.rc, s.it = [0.0, 0.0], 0, 0, 0

    def vec(mode, it, a=None, b_=None, c=None, d=None, e=None, f=None):
        if mode == "HALF":
            if not (s.done == 2 and s.it == it):
                return
        elif s.done:
            return
        al, om, be = s.alpha, s.omega, s.beta
        if mode == "DOT2":
            s.acc[0] += a @ b_
            if c is not None:
                s.acc[1] += c @ d
        elif mode == "S":
            a[:] = b_ - al * c
            s.acc[0] += a @ a
        elif mode == "XR":
            a += al * b_ + om * c
            d[:] = c - om * e
            s.acc[0] += d @ d
            s.acc[1] += f @ d
        elif mode == "P":
            a[:] = b_ + be * (a - om * c)
        elif mode == "HALF":
            a += al * b_
        elif mode == "CG_XR":
            a += al * b_
            c -= al * d
        elif mode == "CG_P":
            a[:] = b_ + be * a
        elif mode == "COPY2":
            a[:] = b_
            if c is not None:
                c[:] = b_

    def scal(step):
        if s.done:
            return
        a0, a1 = s.acc
        s.acc = [0.0, 0.0]
        if step == "BN":
            s.bn = np.sqrt(a0)
            s.tol = max(rtol * s.bn, atol)
        elif step in ("R0", "CG_R0"):
            s.rn = np.sqrt(a0)
            if s.rn <= s.tol:
                s.done, s.rc = 1, 1
            elif s.rn != s.rn:
                s.done, s.rc = 1, -1
            elif maxit == 0:
                s.done, s.rc = 1, 0
            s.rho = a0
            s.it = 0 if s.done else 1
        elif step == "CG_GAMMA0":
            s.rho = a0
        elif step == "ALPHA":
            if a0 == 0:
                s.done, s.rc = 1, -1
                return
            s.alpha = s.rho / a0
        elif step == "SN":
            if np.sqrt(a0) <= s.tol:
                s.rn, s.done, s.rc = np.sqrt(a0), 2, 1
        elif step == "OMEGA":
            if a0 == 0:
                s.done, s.rc = 1, -1
                return
            s.omega = a1 / a0
        elif step in ("END", "CG_RN"):
            s.rn = np.sqrt(a0)
            if s.rn <= s.tol:
                s.done, s.rc = 1, 1
                return
            if s.rn != s.rn:
                s.done, s.rc = 1, -1
                return
            s.beta = (a1 / s.rho) * (s.alpha / s.omega) if step == "END" else a1 / s.rho
            s.rho = a1
            if s.it >= maxit:
                s.done, s.rc = 1, 0
                return
            s.it += 1
        elif step == "CG_ALPHA":
            if not a0 > 0:
                s.done, s.rc = 1, -1
                return
            s.alpha = s.rho / a0

    def P(src, dst):
        if not s.done:
            dst[:] = prec(src)

    def spmv(xin, out, sub):
        if not s.done:
            out[:] = (b - A @ xin) if sub else A @ xin

    r, rh, p, v, ss, t, tmp, z = (np.zeros(n) for _ in range(8))
    P(b, tmp); vec("DOT2", 0, tmp, tmp); scal("BN"); spmv(x, tmp, 1)  # noqa: E702
    if method == "bicgstab":
        P(tmp, r); vec("DOT2", 0, r, r); scal("R0"); vec("COPY2", 0, rh, r, p)  # noqa: E702
    else:
        vec("COPY2", 0, r, tmp); P(r, z); vec("DOT2", 0, z, z); scal("CG_R0")  # noqa: E702
        vec("DOT2", 0, r, z); scal("CG_GAMMA0"); vec("COPY2", 0, p, z)  # noqa: E702
    it = 1
    while not (s.done or it > maxit):
        for _ in range(chunk):
            if it > maxit:
                break
            if method == "bicgstab":
                spmv(p, tmp, 0); P(tmp, v); vec("DOT2", it, rh, v); scal("ALPHA"); vec("S", it, ss, r, v); scal("SN")  # noqa: E702
                vec("HALF", it, x, p); spmv(ss, tmp, 0); P(tmp, t); vec("DOT2", it, t, t, t, ss); scal("OMEGA")  # noqa: E702
                vec("XR", it, x, p, ss, r, t, rh); scal("END"); vec("P", it, p, r, v)  # noqa: E702
            else:
                spmv(p, tmp, 0); vec("DOT2", it, p, tmp); scal("CG_ALPHA"); vec("CG_XR", it, x, p, r, tmp); P(r, z)  # noqa: E702
                vec("DOT2", it, z, z, r, z); scal("CG_RN"); vec("CG_P", it, p, z)  # noqa: E702
            it += 1
    return s


@pytest.mark.parametrize("method", ["bicgstab", "cg"])
def test_model_of_the_device_krylov_schedule_follows_the_oracle(method):
    n = 250
    A = _fe_like(n, 0.03, 1)
    if method == "cg":
        A = (A @ A.T + sp.eye(n)).tocsr()
        A.sort_indices()
    rc, rp, ci, va, lu, dptr = _ilu0(A)
    prec = lambda r: _apply(rp, ci, lu, dptr, r)  # noqa: E731
    rng = np.random.default_rng(2)
    b = rng.standard_normal(n)
    for rtol in (1e-5, 1e-10):
        for maxit in (10000, 3, 0):
            x0 = 0.1 * rng.standard_normal(n)
            x = x0.copy()
            s = _device_schedule(A, prec, method, b, x, rtol, 1e-50, maxit)
            xo = x0.copy()
            try:
                it_o, rn_o, bn_o = fem.solve_krylov_c(A, xo, b, method, "ilu", rtol=rtol, maxiter=maxit)
                assert s.rc == 1 and s.it == it_o and abs(s.rn - rn_o) <= 1e-9 * bn_o and abs(s.bn - bn_o) <= 1e-13 * bn_o
                assert np.max(np.abs(x - xo)) <= 1e-12 * np.max(np.abs(xo))
            except RuntimeError as e:
                assert s.rc == 0 and s.it == maxit and ("after %d iterations" % maxit) in str(e)
