"""Mesh provider: what the drop-in needs from dolfin's mesh layer (vertex coordinates, UFC-ordered
cells, exterior facets with markers).  When dolfin is importable use `from_dolfin(mesh)`; otherwise the
builders below generate the reference's own structured meshes:

  RectangleMesh       dolfin RectangleMesh(Point,Point,nx,ny), default "right" diagonal
  Skew_Mesh           lid-driven-cavity/fenics_fem.py:91-140   (y -> cos(pi y/2))
  DGP_structured_mesh lid-driven-cavity/fenics_fem.py:142-178  (x -> (1-cos(pi x))/2)
  annulus_mesh        structured O-grid stand-in for JBP_mesh (mshr, fenics_base.py:70-91)
"""
import numpy as np

pi = 3.14159265359  # the reference's helper modules override pi (fenics_fem.py:20)
DOLFIN_EPS = 3.0e-16

_FACET_VERTS = np.array([[1, 2], [0, 2], [0, 1]])  # local facet i is opposite local vertex i


class Mesh:
    def __init__(self, coordinates, cells, boundary_facets=None):
        """boundary_facets = (cell, local facet, marker) arrays overrides the exterior-facet search: a
        partition's sub-mesh has artificial exterior facets along the cut that are NOT domain boundary."""
        self._x = np.ascontiguousarray(coordinates, dtype=np.float64)
        c = np.sort(np.asarray(cells, dtype=np.int64), axis=1)  # UFC: ascending vertex ids
        self._cells = np.ascontiguousarray(c.astype(np.int32))
        if boundary_facets is None:
            self._find_exterior_facets()
            self.bfacet_marker = np.zeros(len(self.bfacet_cell), dtype=np.int32)
        else:
            bc, bl, bm = boundary_facets
            self.bfacet_cell = np.ascontiguousarray(bc, dtype=np.int32)
            self.bfacet_local = np.ascontiguousarray(bl, dtype=np.int32)
            self.bfacet_marker = np.ascontiguousarray(bm, dtype=np.int32)
            fv = _FACET_VERTS[self.bfacet_local]
            self.bfacet_verts = np.take_along_axis(c[self.bfacet_cell], fv, axis=1)
            self.bfacet_mid = self._x[self.bfacet_verts].mean(axis=1) if len(bc) else np.zeros((0, 2))

    # -- dolfin-like accessors ------------------------------------------------------------------
    def coordinates(self):
        return self._x

    def cells(self):
        return self._cells

    def num_cells(self):
        return len(self._cells)

    def num_vertices(self):
        return len(self._x)

    def _edge_lengths(self):
        x = self._x[self._cells]
        return np.stack([np.linalg.norm(x[:, a] - x[:, b], axis=1) for a, b in _FACET_VERTS], 1)

    def hmin(self):
        return float(self._edge_lengths().max(axis=1).min())

    def hmax(self):
        return float(self._edge_lengths().max(axis=1).max())

    # -- exterior facets ---------------------------------------------------------------------------
    def _find_exterior_facets(self):
        c = self._cells.astype(np.int64)
        nv = len(self._x)
        key = np.concatenate([c[:, a] * nv + c[:, b] for a, b in _FACET_VERTS])
        cell = np.tile(np.arange(len(c)), 3)
        loc = np.repeat(np.arange(3), len(c))
        _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        ext = cnt[inv] == 1
        order = np.lexsort((loc[ext], cell[ext]))
        self.bfacet_cell = cell[ext][order].astype(np.int32)
        self.bfacet_local = loc[ext][order].astype(np.int32)
        fv = _FACET_VERTS[self.bfacet_local]
        self.bfacet_verts = np.take_along_axis(c[self.bfacet_cell], fv, axis=1)
        self.bfacet_mid = self._x[self.bfacet_verts].mean(axis=1)

    def mark_exterior_facets(self, subdomains):
        """subdomains: list of (marker, inside) applied in order like SubDomain.mark(mf, marker) on a
        facet MeshFunction: a facet is marked when its vertices and midpoint are all inside."""
        for marker, inside in subdomains:
            self.bfacet_marker[facets_inside(self, inside)] = marker
        return self.bfacet_marker


def _call_inside(inside, x):
    """evaluate inside(x, on_boundary=True) for an (n,2) array: vectorised callables are used
    directly; dolfin-style SubDomain objects / scalar callables are looped."""
    fn = inside.inside if hasattr(inside, "inside") else inside
    try:
        r = fn(x.T, True)  # x[0], x[1] index coordinate arrays
        r = np.asarray(r)
        if r.shape == (len(x),):
            return r.astype(bool)
    except Exception:
        pass
    return np.array([bool(fn(p, True)) for p in x], dtype=bool)


def facets_inside(mesh, inside):
    xa = mesh.coordinates()[mesh.bfacet_verts[:, 0]]
    xb = mesh.coordinates()[mesh.bfacet_verts[:, 1]]
    return _call_inside(inside, xa) & _call_inside(inside, xb) & _call_inside(inside, mesh.bfacet_mid)


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    if diagonal != "right":
        raise NotImplementedError("only the default 'right' diagonal is generated here")
    nx, ny = int(nx), int(ny)
    gx = np.linspace(p0[0], p1[0], nx + 1)
    gy = np.linspace(p0[1], p1[1], ny + 1)
    xy = np.empty(((nx + 1) * (ny + 1), 2))
    xy[:, 0] = np.tile(gx, ny + 1)
    xy[:, 1] = np.repeat(gy, nx + 1)
    j, i = np.divmod(np.arange(nx * ny), nx)
    ll = j * (nx + 1) + i
    lr, ul, ur = ll + 1, ll + nx + 1, ll + nx + 2
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = np.column_stack([ll, lr, ur])
    cells[1::2] = np.column_stack([ll, ul, ur])
    return Mesh(xy, cells)


def skewcavity(x, y):
    return 0.5 * (1 - np.cos(x * pi)), 0.5 * (1 - np.cos(y * pi))


def xskewcavity(x, y):
    return 0.5 * (1 - np.cos(x * pi)), y


def yskewcavity(x, y):
    return x, np.cos(0.5 * y * pi)


def expskewcavity(x, y, N):
    return 0.5 * (1 + np.tanh(2 * N * (x - 0.5))), 0.5 * (1 + np.tanh(2 * N * (y - 0.5)))


def _mapped(base, fn):
    x = base.coordinates()
    a, b = fn(x[:, 0], x[:, 1])
    return Mesh(np.column_stack([a, b]), base.cells())


def Skew_Mesh(mm, B, L):
    return _mapped(RectangleMesh((0, 0), (B, L), mm * B, mm * L), yskewcavity)


def DGP_structured_mesh(mm, x_0, y_0, x_1, y_1, B, L):
    return _mapped(RectangleMesh((x_0, y_0), (x_1, y_1), mm * B, mm * L), xskewcavity)


def annulus_mesh(n_theta, n_r, x_1=-0.8, y_1=0.0, r_a=1.0, x_2=0.0, y_2=0.0, r_b=2.0):
    """Structured O-grid between the journal (centre (x_1,y_1), radius r_a) and the bearing (centre
    (x_2,y_2), radius r_b): linear blend in r, quads split into two triangles."""
    th = np.linspace(0.0, 2 * np.pi, n_theta, endpoint=False)
    s = np.linspace(0.0, 1.0, n_r + 1)
    inner = np.column_stack([x_1 + r_a * np.cos(th), y_1 + r_a * np.sin(th)])
    outer = np.column_stack([x_2 + r_b * np.cos(th), y_2 + r_b * np.sin(th)])
    xy = (inner[None] * (1 - s)[:, None, None] + outer[None] * s[:, None, None]).reshape(-1, 2)
    k, i = np.divmod(np.arange(n_theta * n_r), n_theta)
    i2 = (i + 1) % n_theta
    a, b = k * n_theta + i, k * n_theta + i2
    c, d = a + n_theta, b + n_theta
    cells = np.empty((2 * n_theta * n_r, 3), dtype=np.int64)
    cells[0::2] = np.column_stack([a, b, d])
    cells[1::2] = np.column_stack([a, c, d])
    return Mesh(xy, cells)


def from_dolfin(dmesh):
    """Adapter for a real dolfin.Mesh (mesh generation stays in dolfin, north_star)."""
    return Mesh(dmesh.coordinates(), dmesh.cells())
