"""ORACLE (test infrastructure, NOT product code) -- numpy/scipy fp64 restatement of what
dolfin 2019.1.0 `assemble / DirichletBC.apply / solve / project / norm` do at the reference's
call sites (e.g. lid-driven-cavity/incomp_LDC.py:328-331, 304-308, 420-421, 434-435).

parity unpinned: see oracle/ufl_lite.py header.  The algorithms restated here (third-party,
un-vendored; versions pinned by fenics-install.sh:28-40):
  * FIAT/FFC 2019.1.0 "default" quadrature and P1/P2/DG0 Lagrange bases
    in UFC local ordering
  * DOLFIN Assembler: cell loop + exterior-facet loop, full cell-block sparsity, ADD_VALUES
  * DOLFIN DirichletBC.apply: identity rows (pattern kept), b[row] = g, columns untouched
  * dolfin.project: consistent mass matrix + exact (LU) solve
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference leg may import this.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.special import roots_jacobi, roots_legendre

from . import ufl_lite as ufl

DOLFIN_EPS = 3.0e-16

# ---- quadrature -----------------------------------------------------------


def _perm3(a, b):
    return [(a, b), (b, a), (b, b)]


def triangle_rule(q):
    """FFC 'default' scheme for a triangle at estimated degree q -> (points (n,2), weights (n,))."""
    if q <= 1:
        return np.array([[1.0 / 3, 1.0 / 3]]), np.array([0.5])
    if q == 2:
        return (np.array([[1.0 / 6, 1.0 / 6], [1.0 / 6, 2.0 / 3], [2.0 / 3, 1.0 / 6]]),
                np.full(3, 1.0 / 6))
    if q == 3:
        a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
        return (np.array([[a, b], [a, c], [b, a], [b, c], [c, a], [c, b]]), np.full(6, 1.0 / 12))
    if q == 4:
        a, b = 0.816847572980459, 0.091576213509771
        c, d = 0.108103018168070, 0.445948490915965
        p = [(a, b), (b, a), (b, b), (c, d), (d, c), (d, d)]
        w = [0.109951743655322 / 2] * 3 + [0.223381589678011 / 2] * 3
        return np.array(p), np.array(w)
    if q == 5:
        p = [(1.0 / 3, 1.0 / 3)] + _perm3(0.79742698535308720, 0.10128650732345633) \
            + _perm3(0.05971587178976981, 0.47014206410511505)
        w = [0.225 / 2] + [0.12593918054482717 / 2] * 3 + [0.13239415278850616 / 2] * 3
        return np.array(p), np.array(w)
    if q == 6:
        a, b, c = 0.636502499121399, 0.310352451033785, 0.053145049844816
        p = _perm3(0.873821971016996, 0.063089014491502) + _perm3(0.501426509658179, 0.249286745170910) \
            + [(a, b), (b, a), (a, c), (c, a), (b, c), (c, b)]
        w = [0.050844906370207 / 2] * 3 + [0.116786275726379 / 2] * 3 + [0.082851075618374 / 2] * 6
        return np.array(p), np.array(w)
    m = (q + 2) // 2  # collapsed Gauss-Jacobi, FIAT CollapsedQuadratureTriangleRule
    xa, wa = roots_legendre(m)
    xb, wb = roots_jacobi(m, 1.0, 0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            pts.append((0.25 * (1 + xa[i]) * (1 - xb[j]), 0.5 * (1 + xb[j])))
            wts.append(wa[i] * wb[j] / 8.0)
    return np.array(pts), np.array(wts)


def interval_rule(q):
    """Gauss-Legendre on [0,1] with m=(q+2)//2 points (FIAT make_quadrature on an interval)."""
    m = (q + 2) // 2
    x, w = roots_legendre(m)
    return 0.5 * (x + 1.0), 0.5 * w


# ---- reference elements (UFC ordering) -----------------------------------------------------------
NB = {"P2": 6, "P1": 3, "DG0": 1}
DEG = {"P2": 2, "P1": 1, "DG0": 0}


def tabulate(kind, X):
    """X (...,2) reference points -> phi (...,nb), dphi (...,nb,2) (reference derivatives)."""
    xi, et = X[..., 0], X[..., 1]
    l0, l1, l2 = 1.0 - xi - et, xi, et
    one, zero = np.ones_like(xi), np.zeros_like(xi)
    if kind == "DG0":
        return one[..., None], np.zeros(xi.shape + (1, 2))
    if kind == "P1":
        phi = np.stack([l0, l1, l2], -1)
        d = np.stack([np.stack([-one, -one], -1), np.stack([one, zero], -1), np.stack([zero, one], -1)], -2)
        return phi, d
    if kind == "P2":
        phi = np.stack([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1),
                        4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1], -1)
        d = np.stack([
            np.stack([-(4 * l0 - 1), -(4 * l0 - 1)], -1),
            np.stack([4 * l1 - 1, zero], -1),
            np.stack([zero, 4 * l2 - 1], -1),
            np.stack([4 * l2, 4 * l1], -1),
            np.stack([-4 * l2, 4 * (l0 - l2)], -1),
            np.stack([4 * (l0 - l1), -4 * l1], -1)], -2)
        return phi, d
    raise ValueError(kind)


def tabulate_hessian(kind):
    """constant reference second derivatives d2 phi_b / d xi_a d xi_c, shape (nb,2,2); zero for P1 / DG0.
    P2 (UFC order): phi_i = l_i (2 l_i - 1), phi_3 = 4 l1 l2, phi_4 = 4 l0 l2, phi_5 = 4 l0 l1 with l0 = 1 - xi - eta."""
    if kind != "P2":
        return np.zeros((NB[kind], 2, 2))
    dl = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])  # grad of l0, l1, l2
    H = np.zeros((6, 2, 2))
    for i in range(3):
        H[i] = 4.0 * np.outer(dl[i], dl[i])
    for b, (i, j) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
        H[b] = 4.0 * (np.outer(dl[i], dl[j]) + np.outer(dl[j], dl[i]))
    return H


# reference coordinates of the nodes (dof points) per element kind
NODES = {"P1": np.array([[0.0, 0], [1, 0], [0, 1]]),
         "P2": np.array([[0.0, 0], [1, 0], [0, 1], [0.5, 0.5], [0, 0.5], [0.5, 0]]),
         "DG0": np.array([[1.0 / 3, 1.0 / 3]])}


# ---- mesh ----------------------------------------------------------------------------------------
class Mesh:
    """coords (nv,2) f64; cells (nc,3) int, each row ascending (UFC ordering)."""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=float)
        cells = np.asarray(cells, dtype=np.int64)
        if np.any(np.diff(cells, axis=1) <= 0):
            raise ValueError("cell vertices must be sorted ascending (UFC convention)")
        self.cells = cells
        x = self.coords[cells]  # (nc,3,2)
        self.x0 = x[:, 0]
        J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=-1)  # J[:, i, j] = dx_i/dxi_j
        self.J = J
        self.detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        K = np.empty_like(J)  # K = J^-1, K[:, a, i] = dxi_a/dx_i
        K[:, 0, 0], K[:, 0, 1] = J[:, 1, 1] / self.detJ, -J[:, 0, 1] / self.detJ
        K[:, 1, 0], K[:, 1, 1] = -J[:, 1, 0] / self.detJ, J[:, 0, 0] / self.detJ
        self.K = K
        e = [np.linalg.norm(x[:, a] - x[:, b], axis=1) for a, b in ((1, 2), (0, 2), (0, 1))]
        self.h = np.maximum(np.maximum(e[0], e[1]), e[2])  # CellDiameter: max vertex distance
        self.facet_len = np.stack(e, axis=1)  # local facet i is opposite local vertex i
        self._boundary()

    def num_cells(self):
        return len(self.cells)

    def num_vertices(self):
        return len(self.coords)

    def hmin(self):
        return self.h.min()

    def _boundary(self):
        c = self.cells
        loc = [(1, 2), (0, 2), (0, 1)]
        keys, owner = [], []
        nv = len(self.coords)
        for f, (a, b) in enumerate(loc):
            keys.append(c[:, a] * nv + c[:, b])  # sorted pair since rows ascending
            owner.append(np.stack([np.arange(len(c)), np.full(len(c), f)], 1))
        keys = np.concatenate(keys)
        owner = np.concatenate(owner)
        _, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
        bmask = cnt[inv] == 1
        bf = owner[bmask]
        order = np.lexsort((bf[:, 1], bf[:, 0]))
        self.bfacet_cell = bf[order, 0]
        self.bfacet_local = bf[order, 1]
        # outward unit normals + midpoints + vertex ids of the boundary facets
        x = self.coords[c[self.bfacet_cell]]
        ia = np.array([1, 0, 0])[self.bfacet_local]
        ib = np.array([2, 2, 1])[self.bfacet_local]
        r = np.arange(len(ia))
        xa, xb, xo = x[r, ia], x[r, ib], x[r, self.bfacet_local]
        t = xb - xa
        n = np.stack([t[:, 1], -t[:, 0]], 1)
        n /= np.linalg.norm(n, axis=1)[:, None]
        flip = np.einsum("ij,ij->i", n, xo - xa) > 0
        n[flip] *= -1
        self.bfacet_normal = n
        self.bfacet_mid = 0.5 * (xa + xb)
        self.bfacet_verts = np.stack([c[self.bfacet_cell, ia], c[self.bfacet_cell, ib]], 1)
        self.bfacet_marker = np.zeros(len(r), dtype=np.int64)

    def mark_boundary(self, rules, default=0):
        """rules: list of (marker, predicate(x_mid (n,2)) -> bool mask), applied in order
        (later marks overwrite, like successive SubDomain.mark calls)."""
        m = np.full(len(self.bfacet_cell), default, dtype=np.int64)
        for marker, pred in rules:
            m[pred(self.bfacet_mid)] = marker
        self.bfacet_marker = m
        return m


def rectangle_mesh(x0, y0, x1, y1, nx, ny, mapping=None):
    """dolfin RectangleMesh(Point(x0,y0),Point(x1,y1),nx,ny) with the default 'right' diagonal
    (SURVEY.md Appendix B 'Other facts'), optionally followed by a vertex map (x,y)->(x',y') as the
    reference's Skew_Mesh does (lid-driven-cavity/fenics_fem.py:91-140)."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys)
    coords = np.stack([X.ravel(), Y.ravel()], 1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny))
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2 = v0 + 1, v0 + nx + 1
    v3 = v2 + 1
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = np.stack([v0, v1, v3], 1)
    cells[1::2] = np.stack([v0, v2, v3], 1)
    if mapping is not None:
        a, b = mapping(coords[:, 0], coords[:, 1])
        coords = np.stack([a, b], 1)
    return Mesh(coords, cells)


PI_REF = 3.14159265359  # the reference overrides math.pi (lid-driven-cavity/fenics_fem.py:20)


def yskewcavity(x, y):  # lid-driven-cavity/fenics_fem.py:70-73
    return x, np.cos(0.5 * y * PI_REF)


def xskewcavity(x, y):  # lid-driven-cavity/fenics_fem.py:65-68
    return 0.5 * (1 - np.cos(x * PI_REF)), y


def skew_mesh(mm, B=1, L=1):
    """Skew_Mesh(mm,B,L) (lid-driven-cavity/fenics_fem.py:91-140) without refine_top."""
    return rectangle_mesh(0, 0, B, L, mm * B, mm * L, yskewcavity)


# ---- function spaces -----------------------------------------------------------------------------
class FunctionSpace:
    """comps: list of element kinds, one per scalar component (component-blocked local dofs);
    cell_dofs (nc, ld) global dof numbers (ANY numbering, e.g. dolfin's)."""

    def __init__(self, mesh, comps, cell_dofs=None, name=""):
        self.mesh, self.comps, self.name = mesh, list(comps), name
        self.offs = np.concatenate([[0], np.cumsum([NB[k] for k in comps])]).astype(int)
        self.ld = int(self.offs[-1])
        if cell_dofs is None:
            cell_dofs = default_cell_dofs(mesh, comps)
        self.cell_dofs = np.asarray(cell_dofs, dtype=np.int64)
        assert self.cell_dofs.shape == (mesh.num_cells(), self.ld)
        self.dim = int(self.cell_dofs.max()) + 1
        self.scalar = len(comps) == 1

    def degree(self):
        return max(DEG[k] for k in self.comps)

    def comp_degree(self, c):
        return DEG[self.comps[c]]

    def ncomp(self):
        return len(self.comps)

    def tabulate_dof_coordinates(self):
        out = np.zeros((self.dim, 2))
        for c, kind in enumerate(self.comps):
            ref = NODES[kind]
            xy = self.mesh.x0[:, None, :] + np.einsum("cij,nj->cni", self.mesh.J, ref)
            out[self.cell_dofs[:, self.offs[c]:self.offs[c + 1]].ravel()] = xy.reshape(-1, 2)
        return out

    def dof_component(self):
        comp = np.zeros(self.dim, dtype=np.int64)
        for c in range(len(self.comps)):
            comp[self.cell_dofs[:, self.offs[c]:self.offs[c + 1]].ravel()] = c
        return comp


def mesh_edges(mesh):
    c = mesh.cells
    nv = mesh.num_vertices()
    loc = [(1, 2), (0, 2), (0, 1)]
    keys = np.stack([c[:, a] * nv + c[:, b] for a, b in loc], 1)  # (nc,3), local edge i opp. vertex i
    uniq, inv = np.unique(keys.ravel(), return_inverse=True)
    return inv.reshape(-1, 3), len(uniq)


def default_cell_dofs(mesh, comps):
    """Oracle's own simple numbering: component-blocked, vertices then edges (P2), cells (DG0)."""
    nc, nv = mesh.num_cells(), mesh.num_vertices()
    cell_edges, ne = mesh_edges(mesh)
    cols, base = [], 0
    for kind in comps:
        if kind == "P1":
            cols.append(base + mesh.cells)
            base += nv
        elif kind == "P2":
            cols.append(np.concatenate([base + mesh.cells, base + nv + cell_edges], 1))
            base += nv + ne
        else:
            cols.append(base + np.arange(nc)[:, None])
            base += nc
    return np.concatenate(cols, 1)


class Function:
    def __init__(self, space, vec=None):
        self.space = space
        self.vec = np.zeros(space.dim) if vec is None else np.array(vec, dtype=float)

    def expr(self, comps=None):
        if comps is None:
            comps = range(self.space.ncomp())
            return ufl.Coef(self, comps, self.space.scalar)
        return ufl.Coef(self, comps, False)

    def split_comps(self, groups):
        """views of component groups, e.g. W: [(0,1),(2,3,4)] -> (u, D_vec)."""
        return tuple(ufl.Coef(self, g, False) for g in groups)

    def assign(self, other):
        self.vec[:] = other.vec

    def copy(self):
        return Function(self.space, self.vec.copy())


def TrialFunction(space, comps=None):
    return _arg(space, 1, comps)


def TestFunction(space, comps=None):
    return _arg(space, 0, comps)


def _arg(space, number, comps):
    if comps is None:
        return ufl.Arg(space, number, range(space.ncomp()), space.scalar)
    return ufl.Arg(space, number, comps, False)


# ---- evaluation context --------------------------------------------------------------------------
class Ctx:
    def __init__(self, mesh, ents, X, normal=None):
        self.mesh, self.ents, self.X, self.normal = mesh, ents, X, normal
        self.h = mesh.h[ents]
        self.memo, self.keep, self._tab, self.args = {}, [], {}, {}

    def basis(self, kind):
        r = self._tab.get(kind)
        if r is None:
            phi, dref = tabulate(kind, self.X)  # (ne,nq,nb), (ne,nq,nb,2)
            K = self.mesh.K[self.ents]  # (ne, a, i)
            dphys = np.einsum("eqba,eai->eqbi", dref, K)
            r = (phi, dphys)
            self._tab[kind] = r
        return r

    def _coef(self, fn, comps):
        key = ("c", id(fn), comps)
        r = self.memo.get(key)
        if r is None:
            sp_ = fn.space
            vals, grads = [], []
            for c in comps:
                kind = sp_.comps[c]
                phi, dphi = self.basis(kind)
                d = fn.vec[sp_.cell_dofs[self.ents, sp_.offs[c]:sp_.offs[c + 1]]]  # (ne,nb)
                vals.append(np.einsum("eqb,eb->eq", phi, d))
                grads.append(np.einsum("eqbi,eb->eqi", dphi, d))
            r = (np.stack(vals, -1), np.stack(grads, -2))
            self.memo[key] = r
            self.keep.append(fn)
        return r

    def coef_value(self, fn, comps):
        return self._coef(fn, comps)[0]

    def coef_grad(self, fn, comps):
        return self._coef(fn, comps)[1]

    def coef_hess(self, fn, comps):
        """second derivatives of a Function's components, (ne,nq,ncomp,2,2) [.., i, j] = d2 f / dx_i dx_j: constant per
        cell on affine triangles (K^T H_ref K), zero for P1 / DG0 components.  Needed by forms that differentiate a
        function of grad(u), e.g. div(phi_def(u0) (...)) of fenepmp_jbp.py:519 with lambda_D != 0."""
        key = ("h", id(fn), comps)
        r = self.memo.get(key)
        if r is None:
            sp_ = fn.space
            K = self.mesh.K[self.ents]  # (ne, a, i)
            ne, nq = self.X.shape[:2]
            out = []
            for c in comps:
                Href = tabulate_hessian(sp_.comps[c])  # (nb, a, c)
                d = fn.vec[sp_.cell_dofs[self.ents, sp_.offs[c]:sp_.offs[c + 1]]]  # (ne,nb)
                hc = np.einsum("eb,bac,eai,ecj->eij", d, Href, K, K)
                out.append(np.broadcast_to(hc[:, None], (ne, nq, 2, 2)))
            r = np.stack(out, -3)
            self.memo[key] = r
            self.keep.append(fn)
        return r


def _slots(space):
    """list of (component, kind) with kind 0=value, 1=d/dx, 2=d/dy; DG0 has only the value slot."""
    out = []
    for c, k in enumerate(space.comps):
        out.append((c, 0))
        if k != "DG0":
            out += [(c, 1), (c, 2)]
    return out


def _slot_basis(ctx, space):
    """B[e,q,a,s]: value of local basis function a in slot s."""
    sl = _slots(space)
    ne, nq = ctx.X.shape[:2]
    B = np.zeros((ne, nq, space.ld, len(sl)))
    for s, (c, k) in enumerate(sl):
        phi, dphi = ctx.basis(space.comps[c])
        a0, a1 = space.offs[c], space.offs[c + 1]
        B[:, :, a0:a1, s] = phi if k == 0 else dphi[..., k - 1]
    return B


def _set_arg(ctx, number, space, slot):
    val = np.zeros(space.ncomp())
    grad = np.zeros((space.ncomp(), 2))
    if slot is not None:
        c, k = slot
        if k == 0:
            val[c] = 1.0
        else:
            grad[c, k - 1] = 1.0
    ctx.args[number] = (val, grad)


def _find_arg_space(expr, number, seen=None):
    if isinstance(expr, ufl.Arg):
        return expr.space if expr.number == number else None
    for name in ("a", "b"):
        ch = getattr(expr, name, None)
        if isinstance(ch, ufl.Expr) and number in ch.deps:
            r = _find_arg_space(ch, number)
            if r is not None:
                return r
    for it in getattr(expr, "items", []):
        if number in it.deps:
            r = _find_arg_space(it, number)
            if r is not None:
                return r
    return None


def _integral_groups(form, mesh):
    """Merge integrals the way UFL does (one integrand per (type, subdomain); the 'everywhere'
    ds integrand is appended to every ds(i)) and pick the quadrature degree of the merged sum.
    Returns list of (kind, entity selector, integrand, degree)."""
    cell_terms = [i.integrand for i in form.integrals if i.kind == "dx"]
    out = []
    if cell_terms:
        e = cell_terms[0]
        for t in cell_terms[1:]:
            e = e + t
        out.append(("dx", None, e, e.deg()))
    everywhere = [i.integrand for i in form.integrals if i.kind == "ds" and i.subdomain_id is None]
    ids = sorted({i.subdomain_id for i in form.integrals if i.kind == "ds" and i.subdomain_id is not None})
    for sid in ids:
        terms = [i.integrand for i in form.integrals if i.kind == "ds" and i.subdomain_id == sid] + everywhere
        e = terms[0]
        for t in terms[1:]:
            e = e + t
        out.append(("ds", ("in", [sid]), e, e.deg()))
    if everywhere:
        e = everywhere[0]
        for t in everywhere[1:]:
            e = e + t
        out.append(("ds", ("notin", ids), e, e.deg()))
    return out


def _make_ctx(mesh, kind, sel, degree):
    if kind == "dx":
        pts, wts = triangle_rule(degree)
        ents = np.arange(mesh.num_cells())
        X = np.broadcast_to(pts[None], (len(ents),) + pts.shape)
        W = wts[None, :] * np.abs(mesh.detJ)[:, None]
        return Ctx(mesh, ents, X), W
    s, w = interval_rule(degree)
    mk = mesh.bfacet_marker
    mask = np.isin(mk, sel[1]) if sel[0] == "in" else ~np.isin(mk, sel[1])
    cells, lf = mesh.bfacet_cell[mask], mesh.bfacet_local[mask]
    ref = NODES["P1"]
    ia = np.array([1, 0, 0])[lf]
    ib = np.array([2, 2, 1])[lf]
    X = ref[ia][:, None, :] * (1 - s)[None, :, None] + ref[ib][:, None, :] * s[None, :, None]
    W = w[None, :] * mesh.facet_len[cells, lf][:, None]
    return Ctx(mesh, cells, X, mesh.bfacet_normal[mask]), W


def quadrature_degrees(form, mesh=None):
    """degree per merged integral -- exposed so the CUDA side can be checked to use the same rule."""
    return [(k, sel, d) for k, sel, _, d in _integral_groups(form, mesh)]


def assemble(form, mesh=None):
    """dolfin assemble(form): rank 0 -> float, rank 1 -> ndarray, rank 2 -> scipy CSR with the
    full cell-block pattern (explicit zeros kept, columns sorted)."""
    rank = form.rank()
    tspace = rspace = None
    for i in form.integrals:
        if tspace is None and 0 in i.integrand.deps:
            tspace = _find_arg_space(i.integrand, 0)
        if rspace is None and 1 in i.integrand.deps:
            rspace = _find_arg_space(i.integrand, 1)
    if mesh is None:
        mesh = (tspace or rspace).mesh if (tspace or rspace) is not None else None
    if mesh is None:
        mesh = _find_mesh(form)
    if rank == 0:
        total = 0.0
        for kind, sel, e, d in _integral_groups(form, mesh):
            ctx, W = _make_ctx(mesh, kind, sel, d)
            v = np.broadcast_to(e.ev(ctx), W.shape)
            total += float(np.sum(W * v))
        return total
    if rank == 1:
        b = np.zeros(tspace.dim)
        for kind, sel, e, d in _integral_groups(form, mesh):
            ctx, W = _make_ctx(mesh, kind, sel, d)
            if len(ctx.ents) == 0:
                continue
            B = _slot_basis(ctx, tspace)
            be = np.zeros((len(ctx.ents), tspace.ld))
            for s, slot in enumerate(_slots(tspace)):
                _set_arg(ctx, 0, tspace, slot)
                F = np.broadcast_to(e.ev(ctx), W.shape)
                be += np.einsum("eq,eqa->ea", W * F, B[..., s])
            np.add.at(b, tspace.cell_dofs[ctx.ents].ravel(), be.ravel())
        return b
    # rank 2
    rows_all, cols_all, vals_all = [], [], []
    # full cell-block pattern over ALL cells (SparsityPatternBuilder), explicit zeros
    cd_t, cd_r = tspace.cell_dofs, rspace.cell_dofs
    rows_all.append(np.repeat(cd_t, rspace.ld, axis=1).ravel())
    cols_all.append(np.tile(cd_r, (1, tspace.ld)).ravel())
    vals_all.append(np.zeros(cd_t.shape[0] * tspace.ld * rspace.ld))
    for kind, sel, e, d in _integral_groups(form, mesh):
        ctx, W = _make_ctx(mesh, kind, sel, d)
        if len(ctx.ents) == 0:
            continue
        Bt, Br = _slot_basis(ctx, tspace), _slot_basis(ctx, rspace)
        Ae = np.zeros((len(ctx.ents), tspace.ld, rspace.ld))
        for t, tslot in enumerate(_slots(rspace)):
            _set_arg(ctx, 1, rspace, tslot)
            WB = W[..., None] * Br[..., t]  # (e,q,b)
            for s, sslot in enumerate(_slots(tspace)):
                _set_arg(ctx, 0, tspace, sslot)
                C = e.ev(ctx)
                if not np.any(C):
                    continue
                C = np.broadcast_to(C, W.shape)
                Ae += np.einsum("eqa,eqb->eab", C[..., None] * Bt[..., s], WB)
        rows_all.append(np.repeat(cd_t[ctx.ents], rspace.ld, axis=1).ravel())
        cols_all.append(np.tile(cd_r[ctx.ents], (1, tspace.ld)).ravel())
        vals_all.append(Ae.ravel())
    A = sp.coo_matrix((np.concatenate(vals_all), (np.concatenate(rows_all), np.concatenate(cols_all))),
                      shape=(tspace.dim, rspace.dim)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def _find_mesh(form):
    def walk(e):
        if isinstance(e, ufl.Coef):
            return e.fn.space.mesh
        for name in ("a", "b"):
            ch = getattr(e, name, None)
            if isinstance(ch, ufl.Expr):
                r = walk(ch)
                if r is not None:
                    return r
        for it in getattr(e, "items", []):
            r = walk(it)
            if r is not None:
                return r
        return None
    for i in form.integrals:
        r = walk(i.integrand)
        if r is not None:
            return r
    raise ValueError("cannot infer mesh of a form without coefficients/arguments")


# ---- Dirichlet conditions ------------------------------------------------------------------------
class DirichletBC:
    """DirichletBC(V.sub(...), g, sub_domain) with method 'topological': dofs in the closure of
    boundary facets whose two vertices AND midpoint satisfy inside(x, on_boundary=True); values =
    g at the dof coordinates.  `comps` selects the sub-space components (W.sub(0) -> (0,1))."""

    def __init__(self, space, comps, value, inside):
        self.space = space
        mesh = space.mesh
        xa = mesh.coords[mesh.bfacet_verts[:, 0]]
        xb = mesh.coords[mesh.bfacet_verts[:, 1]]
        ok = inside(xa) & inside(xb) & inside(mesh.bfacet_mid)
        cells, lf = mesh.bfacet_cell[ok], mesh.bfacet_local[ok]
        dofs = []
        for c in comps:
            kind = space.comps[c]
            o = space.offs[c]
            if kind == "P1":
                loc = np.array([[1, 2], [0, 2], [0, 1]])[lf]
            elif kind == "P2":
                loc = np.array([[1, 2, 3], [0, 2, 4], [0, 1, 5]])[lf]
            else:
                continue
            dofs.append((space.cell_dofs[cells[:, None], o + loc].ravel(),
                         np.full(loc.size, comps.index(c))))
        d = np.concatenate([x[0] for x in dofs]) if dofs else np.zeros(0, dtype=np.int64)
        ci = np.concatenate([x[1] for x in dofs]) if dofs else np.zeros(0, dtype=np.int64)
        self.dofs, idx = np.unique(d, return_index=True)
        self.dof_comp = ci[idx]
        self.x = space.tabulate_dof_coordinates()[self.dofs]
        self.value = value  # callable (x (n,2)) -> (n, len(comps)) or (n,), or constants

    def values(self):
        v = self.value
        if callable(v):
            v = np.asarray(v(self.x), dtype=float)
        else:
            v = np.broadcast_to(np.asarray(v, dtype=float), (len(self.dofs),) + np.shape(v))
        if v.ndim == 2:
            v = v[np.arange(len(self.dofs)), self.dof_comp]
        return v

    def apply(self, A=None, b=None):
        g = self.values()
        if A is not None:
            apply_identity_rows(A, self.dofs)
        if b is not None:
            b[self.dofs] = g


def apply_identity_rows(A, rows):
    """MatZeroRows(rows, diag=1.0) keeping the pattern (in place on a CSR matrix)."""
    for r in rows:
        s, e = A.indptr[r], A.indptr[r + 1]
        A.data[s:e] = 0.0
        A.data[s:e][A.indices[s:e] == r] = 1.0


# ---- linear algebra ------------------------------------------------------------------------------
def solve_lu(A, b):
    return spla.splu(A.tocsc()).solve(b)


_KLIB = None


def solve_krylov_c(A, x, b, method="bicgstab", pc="ilu", rtol=1e-5, atol=1e-50, maxiter=10000):
    """PETSc-like Krylov solve in C (oracle/krylov_c.c): KSPBCGS / KSPCG with ILU(0) ('ilu', PETSc's serial
    "default"), Jacobi or no preconditioning, left-preconditioned residual test, x = initial guess (the reference sets
    nonzero_initial_guess).  Returns (iterations, ||M^-1 r||, ||M^-1 b||); raises like dolfin on non-convergence."""
    global _KLIB
    import ctypes as C
    if _KLIB is None:
        import importlib.util
        import os
        here = os.path.dirname(os.path.abspath(__file__))
        spec = importlib.util.spec_from_file_location("fx_oracle_build", os.path.join(here, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _KLIB = C.CDLL(mod.build())
        _KLIB.fxo_krylov.restype = C.c_int
    A = A.tocsr()
    A.sort_indices()
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    va = np.ascontiguousarray(A.data, dtype=np.float64)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    xx = np.ascontiguousarray(x, dtype=np.float64).copy()
    out = np.zeros(3)
    vp = C.c_void_p
    rc = _KLIB.fxo_krylov(C.c_int(A.shape[0]), vp(rp.ctypes.data), vp(ci.ctypes.data), vp(va.ctypes.data),
                          vp(xx.ctypes.data), vp(bb.ctypes.data), C.c_int({"bicgstab": 0, "cg": 1}[method]),
                          C.c_int({"none": 0, "jacobi": 1, "ilu": 2, "default": 2}[pc]), C.c_double(rtol),
                          C.c_double(atol), C.c_int(maxiter), vp(out.ctypes.data))
    if rc != 1:
        raise RuntimeError("Krylov solver (%s, %s) did not converge: status %d after %d iterations, residual %.3e"
                           % (method, pc, rc, int(out[0]), out[1]))
    x[:] = xx
    return int(out[0]), float(out[1]), float(out[2])


def solve(A, x, b, method="lu", pc="default", rtol=1e-5, atol=1e-50, maxiter=10000):
    """solve(A, x, b[, method, pc]) writing into the array x (initial guess = contents of x).
    method 'lu' (dolfin default when no method is given) or 'bicgstab'/'cg'/'gmres' with ILU(0)-like
    ('default'), 'jacobi' or 'none' preconditioning.  Oracle parity runs use 'lu'."""
    if method == "lu":
        x[:] = solve_lu(A, b)
        return 0
    if pc in ("default", "ilu"):
        ilu = spla.spilu(A.tocsc(), drop_tol=0.0, fill_factor=1.0)
        M = spla.LinearOperator(A.shape, ilu.solve)
    elif pc == "jacobi":
        dinv = 1.0 / A.diagonal()
        M = spla.LinearOperator(A.shape, lambda v: dinv * v)
    else:
        M = None
    fn = {"bicgstab": spla.bicgstab, "cg": spla.cg, "gmres": spla.gmres}[method]
    sol, info = fn(A, b, x0=x.copy(), rtol=rtol, atol=atol, maxiter=maxiter, M=M)
    if info != 0:
        raise RuntimeError("Krylov solver did not converge (info=%d)" % info)
    x[:] = sol
    return info


def project(expr, space):
    """dolfin.project(v, V): a = inner(w, Pv)*dx, L = inner(w, v)*dx, assemble_system, LU solve
    (dolfin/fem/projection.py).  Returns a new Function (as the reference rebinds names to it)."""
    expr = ufl.as_expr(expr)
    w, Pv = TestFunction(space), TrialFunction(space)
    a = ufl.inner(w, Pv) * ufl.dx
    L = ufl.inner(w, expr) * ufl.dx
    A = assemble(a, space.mesh)
    if not L.integrals:
        return Function(space)
    b = assemble(L, space.mesh)
    return Function(space, solve_lu(A, b))


def norm(vec, kind="l2"):
    if kind == "l2":
        return float(np.sqrt(np.dot(vec, vec)))
    if kind == "linf":
        return float(np.max(np.abs(vec)))
    raise ValueError(kind)


def annulus_mesh(n_theta, n_r, x_1=-0.8, y_1=0.0, r_a=1.0, x_2=0.0, y_2=0.0, r_b=2.0):
    """Structured O-grid stand-in for JBP_mesh (mshr Circle - Circle, fenics_base.py:70-91): journal
    circle (x_1,y_1,r_a) inside bearing circle (x_2,y_2,r_b), linear blend in the radial direction
    (synthetic input S2)."""
    th = 2.0 * np.pi * np.arange(n_theta) / n_theta
    inner_c = np.stack([x_1 + r_a * np.cos(th), y_1 + r_a * np.sin(th)], 1)
    outer_c = np.stack([x_2 + r_b * np.cos(th), y_2 + r_b * np.sin(th)], 1)
    rows = [inner_c + (outer_c - inner_c) * (k / n_r) for k in range(n_r + 1)]
    coords = np.concatenate(rows, 0)
    cells = []
    for k in range(n_r):
        for i in range(n_theta):
            a, b = k * n_theta + i, k * n_theta + (i + 1) % n_theta
            c, d = a + n_theta, b + n_theta
            cells.append(sorted((a, b, d)))
            cells.append(sorted((a, c, d)))
    return Mesh(coords, np.array(cells))
