"""Pins the oracle itself (the reference has no golden vectors -- SURVEY.md 4/8c): construction-
independent checks against exact (sympy) integrals and closed-form identities."""
import numpy as np
import pytest
import sympy as sp

from oracle import fem
from oracle.ufl_lite import (CellDiameter, FacetNormal, as_matrix, div, dot, ds, dx, grad, inner, lhs,
                             nabla_grad, rhs)

X, Y = sp.symbols("x y")
TRIS = [np.array([[0.1, 0.2], [1.3, 0.1], [0.4, 1.1]]),      # positive orientation
        np.array([[0.0, 0.0], [0.2, 0.9], [1.1, 0.3]]),      # negative orientation
        np.array([[2.0, 1.0], [2.001, 1.7], [2.6, 1.05]])]   # thin


def tri_integral(expr, P):
    """exact integral of a sympy expression over triangle P via the affine map from the unit triangle."""
    s, t = sp.symbols("s t")
    x = P[0][0] + (P[1][0] - P[0][0]) * s + (P[2][0] - P[0][0]) * t
    y = P[0][1] + (P[1][1] - P[0][1]) * s + (P[2][1] - P[0][1]) * t
    J = abs((P[1][0] - P[0][0]) * (P[2][1] - P[0][1]) - (P[2][0] - P[0][0]) * (P[1][1] - P[0][1]))
    e = expr.subs({X: x, Y: y}, simultaneous=True)
    return float(sp.integrate(sp.integrate(e, (t, 0, 1 - s)), (s, 0, 1)) * J)


def p2_basis_sym(P):
    """P2 nodal basis on triangle P as sympy polynomials in (x, y), UFC ordering."""
    A = sp.Matrix([[1, P[i][0], P[i][1]] for i in range(3)])
    lam = [sum(c * m for c, m in zip(A.inv()[:, i], (1, X, Y))) for i in range(3)]
    return [lam[0] * (2 * lam[0] - 1), lam[1] * (2 * lam[1] - 1), lam[2] * (2 * lam[2] - 1),
            4 * lam[1] * lam[2], 4 * lam[0] * lam[2], 4 * lam[0] * lam[1]]


def one_cell(P):
    return fem.Mesh(P, np.array([[0, 1, 2]]))


def interp(space, funcs):
    """P2/P1 nodal interpolant of sympy polynomials (exact for degree <= element degree)."""
    f = fem.Function(space)
    xy = space.tabulate_dof_coordinates()
    comp = space.dof_component()
    for i in range(space.dim):
        f.vec[i] = float(funcs[comp[i]].subs({X: xy[i, 0], Y: xy[i, 1]}))
    return f


@pytest.mark.parametrize("q", list(range(0, 16)))
def test_triangle_rules_are_exact_to_their_degree(q):
    pts, w = fem.triangle_rule(q)
    assert abs(w.sum() - 0.5) < 1e-14
    for a in range(q + 1):
        for b in range(q + 1 - a):
            exact = float(sp.factorial(a) * sp.factorial(b) / sp.factorial(a + b + 2))
            assert abs(np.sum(w * pts[:, 0] ** a * pts[:, 1] ** b) - exact) < 1e-13, (q, a, b)


def test_rule_sizes_match_ffc_default_scheme():
    assert [len(fem.triangle_rule(q)[1]) for q in range(0, 16)] == [1, 1, 3, 6, 6, 7, 12, 16, 25, 25, 36, 36, 49, 49, 64, 64]
    assert [len(fem.interval_rule(q)[1]) for q in range(0, 8)] == [1, 1, 2, 2, 3, 3, 4, 4]


@pytest.mark.parametrize("P", TRIS)
def test_p2_mass_and_stiffness_exact(P):
    m = one_cell(P)
    V = fem.FunctionSpace(m, ["P2"])
    u, v = fem.TrialFunction(V), fem.TestFunction(V)
    M = fem.assemble(u * v * dx).toarray()
    K = fem.assemble(inner(grad(u), grad(v)) * dx).toarray()
    phi = p2_basis_sym(P)
    cd = V.cell_dofs[0]
    for a in range(6):
        for b in range(6):
            em = tri_integral(phi[a] * phi[b], P)
            ek = tri_integral(sp.diff(phi[a], X) * sp.diff(phi[b], X) + sp.diff(phi[a], Y) * sp.diff(phi[b], Y), P)
            assert abs(M[cd[a], cd[b]] - em) < 1e-13 * max(1, abs(em))
            assert abs(K[cd[a], cd[b]] - ek) < 1e-11 * max(1, abs(ek))
    assert abs(M.sum() - abs(m.detJ[0]) / 2) < 1e-14


@pytest.mark.parametrize("P", TRIS)
def test_ufl_tensor_semantics_against_index_definitions(P):
    """dot/grad/nabla_grad/div index conventions (UFL 2019.1): checked as exact integrals."""
    m = one_cell(P)
    V2, V3 = fem.FunctionSpace(m, ["P2", "P2"]), fem.FunctionSpace(m, ["P2"] * 3)
    us = [X * X - 2 * X * Y + 0.3, 0.5 * Y * Y + X - 1]
    ts = [1 + X * Y, X * X - Y, 2 * Y * Y + X * Y - X]
    u, t = interp(V2, us).expr(), interp(V3, ts).expr()
    T = as_matrix([[t[0], t[1]], [t[1], t[2]]])
    Ts = [[ts[0], ts[1]], [ts[1], ts[2]]]
    d = (X, Y)
    G = [[sp.diff(us[i], d[j]) for j in range(2)] for i in range(2)]  # grad(u)[i,j] = d u_i / d x_j
    w = [[1.0, -2.0], [0.5, 3.0]]  # arbitrary weights to probe each component
    W = as_matrix(w)
    # dot(u, grad(T))[j,k] = sum_i u_i d_k T_ij   (derivative index LAST, dot contracts u with FIRST index)
    e1 = sum(w[j][k] * sum(us[i] * sp.diff(Ts[i][j], d[k]) for i in range(2)) for j in range(2) for k in range(2))
    assert abs(fem.assemble(inner(dot(u, grad(T)), W) * dx) - tri_integral(e1, P)) < 1e-12
    # dot(u, nabla_grad(T))[i,j] = u_k d_k T_ij
    e2 = sum(w[i][j] * sum(us[k] * sp.diff(Ts[i][j], d[k]) for k in range(2)) for i in range(2) for j in range(2))
    assert abs(fem.assemble(inner(dot(u, nabla_grad(T)), W) * dx) - tri_integral(e2, P)) < 1e-12
    # dot(grad(u), T)[i,k] = G_ij T_jk ; dot(T, grad(u).T)[i,k] = T_ij G_kj
    e3 = sum(w[i][k] * sum(G[i][j] * Ts[j][k] for j in range(2)) for i in range(2) for k in range(2))
    assert abs(fem.assemble(inner(dot(grad(u), T), W) * dx) - tri_integral(e3, P)) < 1e-12
    e4 = sum(w[i][k] * sum(Ts[i][j] * G[k][j] for j in range(2)) for i in range(2) for k in range(2))
    assert abs(fem.assemble(inner(dot(T, grad(u).T), W) * dx) - tri_integral(e4, P)) < 1e-12
    # div(T)_i = d_j T_ij ; div(u) = d_i u_i ; grad(u)*u = G_ij u_j
    e5 = sum((1.0 + i) * sum(sp.diff(Ts[i][j], d[j]) for j in range(2)) for i in range(2))
    assert abs(fem.assemble(inner(div(T), [1.0, 2.0]) * dx) - tri_integral(e5, P)) < 1e-12
    e6 = (G[0][0] + G[1][1]) * ts[0]
    assert abs(fem.assemble(div(u) * t[0] * dx) - tri_integral(e6, P)) < 1e-12
    e7 = sum((1.0 + i) * sum(G[i][j] * us[j] for j in range(2)) for i in range(2))
    assert abs(fem.assemble(inner(grad(u) * u, [1.0, 2.0]) * dx) - tri_integral(e7, P)) < 1e-12
    # nabla_grad(u)*n on a facet: (d_k u_i n_i)_k -- checked through the divergence theorem below


@pytest.mark.parametrize("P", TRIS)
def test_facet_integrals_via_divergence_theorem(P):
    """int_dK f n_k ds = int_K d_k f dx for polynomial f: pins facet quadrature, normals, scaling."""
    m = one_cell(P)
    V1 = fem.FunctionSpace(m, ["P2"])
    f_s = 1 + X * X - 3 * X * Y + Y
    f = interp(V1, [f_s]).expr()
    n = FacetNormal()
    for k, var in enumerate((X, Y)):
        lhs_v = fem.assemble(f * n[k] * ds)
        assert abs(lhs_v - tri_integral(sp.diff(f_s, var), P)) < 1e-12
    # CellDiameter = max vertex distance
    h = max(np.linalg.norm(P[a] - P[b]) for a, b in ((0, 1), (0, 2), (1, 2)))
    assert abs(fem.assemble(CellDiameter() * f * dx) - h * tri_integral(f_s, P)) < 1e-13


def test_lhs_rhs_split_and_zero_folding():
    m = fem.skew_mesh(3)
    Q = fem.FunctionSpace(m, ["P1"])
    p, q = fem.TrialFunction(Q), fem.TestFunction(Q)
    g = fem.Function(Q, np.random.default_rng(0).standard_normal(Q.dim))
    F = (2.0 * (p - g.expr()) / 0.1 + 0.0 * p * p) * q * dx + inner(grad(p) + grad(g.expr()), grad(q)) * dx
    a, L = lhs(F), rhs(F)
    A, b = fem.assemble(a), fem.assemble(L)
    ref_A = fem.assemble((2.0 / 0.1) * p * q * dx + inner(grad(p), grad(q)) * dx)
    ref_b = fem.assemble((2.0 / 0.1) * g.expr() * q * dx - inner(grad(g.expr()), grad(q)) * dx)
    assert abs(A - ref_A).max() < 1e-12 and np.abs(b - ref_b).max() < 1e-12
    # python 0.0 * expr folds to Zero and lowers the estimated degree (UFL FloatValue.__new__)
    assert (0.0 * p * q).deg() == 0 and (1.0 * p * q).deg() == 2


def test_pattern_sizes_match_closed_forms():
    """nnz(W)=346n^2+64n+4, nnz(Zc)=414n^2+144n+9, nnz(Q)=7n^2+6n+1 on the n x n mesh."""
    n = 5
    m = fem.skew_mesh(n)
    for comps, dim, nnz in ((["P2", "P2", "DG0", "DG0", "DG0"], 14 * n * n + 8 * n + 2, 346 * n * n + 64 * n + 4),
                            (["P2"] * 3, 3 * (2 * n + 1) ** 2, 414 * n * n + 144 * n + 9),
                            (["P1"], (n + 1) ** 2, 7 * n * n + 6 * n + 1)):
        V = fem.FunctionSpace(m, comps)
        A = fem.assemble(inner(fem.TrialFunction(V), fem.TestFunction(V)) * dx)
        assert V.dim == dim and A.nnz == nnz


def test_weighted_gauge_solve_of_the_singular_pressure_system():
    """oracle.configs_dgp.solve_weighted_gauge: on a pure-Neumann P1 stiffness matrix (constants in the null space) it
    returns A x = b with sum_i A_ii (x - x0)_i = 0 -- the invariant of LEFT-Jacobi-preconditioned Krylov iterations, checked
    here against an actual Jacobi-preconditioned CG run from the same initial guess."""
    import scipy.sparse.linalg as spla
    from oracle.configs_dgp import solve_weighted_gauge
    from oracle.ufl_lite import dx, grad, inner
    mesh = fem.skew_mesh(6)
    Q = fem.FunctionSpace(mesh, ["P1"])
    p, q = fem.TrialFunction(Q), fem.TestFunction(Q)
    A = fem.assemble(inner(grad(p), grad(q)) * dx).tocsr()
    rng = np.random.default_rng(0)
    b = A @ rng.standard_normal(A.shape[0])          # consistent right-hand side
    x0 = rng.standard_normal(A.shape[0])
    x = solve_weighted_gauge(A, b, x0)
    d = A.diagonal()
    assert np.abs(A @ x - b).max() < 1e-11 * np.abs(b).max()
    assert abs(d @ (x - x0)) < 1e-10 * np.abs(d).max() * np.abs(x).max()
    M = spla.LinearOperator(A.shape, lambda v: v / d)
    xk, info = spla.cg(A, b, x0=x0.copy(), rtol=1e-14, atol=0.0, maxiter=5000, M=M)
    assert info == 0 and np.abs(xk - x).max() < 1e-8 * np.abs(x).max()


def test_second_derivatives_of_functions():
    """oracle support for forms that differentiate a function of grad(u) -- div(phi_def(u0) (f tau0 - I)) of
    fenepmp_jbp.py:519 with lambda_D != 0: (i) the Hessian of an interpolated quadratic is exact; (ii) the product rule
    div(s M) = s div(M) + M grad(s) holds for s = phi_def(u); (iii) the lambda_D != 0 forms assemble with the
    degrees SURVEY.md derived by hand (13, 13, 14)."""
    from oracle.configs_jbp import FenepmpJBP, phi_def
    from oracle.ufl_lite import div, dot, dx, grad, inner
    mesh = fem.skew_mesh(4)
    V = fem.FunctionSpace(mesh, ["P2", "P2"])
    u = fem.Function(V)
    xy, comp = V.tabulate_dof_coordinates(), V.dof_component()
    x, y = xy[:, 0], xy[:, 1]
    u.vec[:] = np.where(comp == 0, 1.5 * x * x - 0.5 * x * y + 2 * y * y + x, -x * x + 3 * x * y + 0.25 * y * y - y)
    Q = fem.FunctionSpace(mesh, ["DG0"])
    q = fem.TestFunction(Q)
    area = fem.assemble(inner(1.0, q) * dx)
    H = grad(grad(u.expr()))
    for (c, i, j), val in {(0, 0, 0): 3.0, (0, 0, 1): -0.5, (0, 1, 1): 4.0, (1, 0, 0): -2.0, (1, 1, 0): 3.0, (1, 1, 1): 0.5}.items():
        assert np.abs(fem.assemble(inner(H[c, i, j], q) * dx) / area - val).max() < 1e-11
    Z = fem.FunctionSpace(mesh, ["P2"] * 3)
    t = fem.Function(Z)
    t.vec[:] = np.random.default_rng(0).standard_normal(Z.dim)
    from oracle.configs import sym
    M = sym(t.expr())
    s = phi_def(u.expr(), 0.3)
    v = fem.TestFunction(V)
    lhs_ = fem.assemble(inner(div(s * M), v) * dx)
    rhs_ = fem.assemble(inner(s * div(M) + dot(M, grad(s)), v) * dx)
    assert np.abs(lhs_ - rhs_).max() < 1e-12 * np.abs(rhs_).max()
    assert np.abs(fem.assemble(inner(dot(M, grad(s)), v) * dx)).max() > 1e-3 * np.abs(rhs_).max()  # the new term carries weight
    cfg = FenepmpJBP(fem.annulus_mesh(8, 2), lambda_d=0.1)
    degs = {n: max(d for _, _, d in fem.quadrature_degrees(cfg.forms[n], cfg.mesh)) for n in ("L2", "a4", "L9")}
    assert degs == {"L2": 13, "a4": 13, "L9": 14}
