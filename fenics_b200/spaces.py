"""Function spaces of the reference's `function_spaces(mesh, order)` (lid-driven-cavity/fenics_fem.py:
262-287, buoyancy-driven-flow/fenics_fem.py:448-468, JBP fenics_base.py:470-502) as cell-dof tables.

With dolfin present the tables must be read from dolfin (`V.dofmap().cell_dofs(c)`, `from_dolfin`) so the
numbering is dolfin's; without it `build_cell_dofs` produces a locality-preserving numbering of its own
(first-touch order of a cell traversal, all components of a mesh entity consecutive).  The CUDA side
accepts ANY numbering -- it only ever sees the tables.
"""
import numpy as np

ELEMENTS = {  # space name -> scalar components (UFC local order is component-blocked)
    "W": ("P2", "P2", "DG0", "DG0", "DG0"),
    "V": ("P2", "P2"),
    "Vs": ("P2",),  # V.sub(0).collapse(): the stream-function space (fenics_fem.py:342)
    "Zc": ("P2", "P2", "P2"),
    "Zd": ("DG0", "DG0", "DG0"),
    "Q": ("P1",),
    "Qt": ("DG0",),
}
_NB = {"P2": 6, "P1": 3, "DG0": 1}
_FACET_VERTS = np.array([[1, 2], [0, 2], [0, 1]])
_P2_NODES = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.5, 0.5], [0.0, 0.5], [0.5, 0.0]])
_P1_NODES = _P2_NODES[:3]


def cell_edges(mesh):
    c = mesh.cells().astype(np.int64)
    nv = mesh.num_vertices()
    key = np.stack([c[:, a] * nv + c[:, b] for a, b in _FACET_VERTS], 1)
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    return inv.reshape(-1, 3), len(uniq)


def build_cell_dofs(mesh, comps):
    """(nc, ld) int32 table for a (mixed) Lagrange/DG0 space with scalar components `comps`."""
    nc, nv = mesh.num_cells(), mesh.num_vertices()
    c = mesh.cells().astype(np.int64)
    n_p2 = sum(k == "P2" for k in comps)
    n_p1 = sum(k == "P1" for k in comps)
    n_dg = sum(k == "DG0" for k in comps)
    ce, ne = (cell_edges(mesh) if n_p2 else (np.zeros((nc, 0), dtype=np.int64), 0))
    # entity ids: vertices [0,nv), edges [nv,nv+ne), cells [nv+ne, nv+ne+nc); block = #dofs on the entity
    block = np.concatenate([np.full(nv, n_p2 + n_p1), np.full(ne, n_p2), np.full(nc, n_dg)])
    ents = np.concatenate([c, nv + ce, (nv + ne + np.arange(nc))[:, None]], axis=1)
    flat = ents.ravel()
    uniq, first = np.unique(flat, return_index=True)
    order = uniq[np.argsort(first, kind="stable")]  # first-touch order
    order = order[block[order] > 0]
    start = np.zeros(len(block), dtype=np.int64)
    start[order] = np.concatenate([[0], np.cumsum(block[order])[:-1]])
    cols = []
    iv = ie = ic = 0  # running index of the component inside vertex / edge / cell blocks
    for k in comps:
        if k == "P2":
            cols.append(np.concatenate([start[c] + iv, start[nv + ce] + ie], axis=1))
            iv += 1
            ie += 1
        elif k == "P1":
            cols.append(start[c] + iv)
            iv += 1
        else:
            cols.append((start[nv + ne + np.arange(nc)] + ic)[:, None])
            ic += 1
    return np.ascontiguousarray(np.concatenate(cols, axis=1).astype(np.int32))


class FunctionSpace:
    def __init__(self, mesh, name, cell_dofs=None, comps=None, parent=None):
        self.mesh, self.name = mesh, name
        self.comps = tuple(comps) if comps is not None else ELEMENTS[name]
        self.parent = parent  # set for W.sub(i) views
        if parent is None:
            self.cell_dofs = build_cell_dofs(mesh, self.comps) if cell_dofs is None else \
                np.ascontiguousarray(cell_dofs, dtype=np.int32)
            self._dim = int(self.cell_dofs.max()) + 1
            self.sub_comps = tuple(range(len(self.comps)))
        self.offsets = np.concatenate([[0], np.cumsum([_NB[k] for k in self.comps])]).astype(int)

    def dim(self):
        return self._dim if self.parent is None else self.parent.dim()

    def ufl_element_degree(self):
        return max({"P2": 2, "P1": 1, "DG0": 0}[k] for k in self.comps)

    def sub(self, i):
        """W.sub(0) = velocity (components 0,1), W.sub(1) = DEVSS tensor (2,3,4); vector spaces: one
        scalar component."""
        root = self.parent or self
        if root.name == "W":
            comps = (0, 1) if i == 0 else (2, 3, 4)
        else:
            comps = (i,)
        s = FunctionSpace(self.mesh, "%s.sub(%d)" % (root.name, i), comps=[root.comps[k] for k in comps], parent=root)
        s.sub_comps = comps
        return s

    def tabulate_dof_coordinates(self):
        root = self.parent or self
        x = self.mesh.coordinates()[self.mesh.cells()]
        out = np.zeros((root.dim(), 2))
        for ci, kind in enumerate(root.comps):
            ref = {"P2": _P2_NODES, "P1": _P1_NODES, "DG0": np.array([[1 / 3, 1 / 3]])}[kind]
            l0 = 1 - ref[:, 0] - ref[:, 1]
            pts = (l0[None, :, None] * x[:, None, 0] + ref[None, :, 0, None] * x[:, None, 1]
                   + ref[None, :, 1, None] * x[:, None, 2])
            out[root.cell_dofs[:, root.offsets[ci]:root.offsets[ci + 1]].ravel()] = pts.reshape(-1, 2)
        return out


def function_spaces(mesh, order=2, dofmaps=None):
    """W, V, Vd, Z, Zd, Zc, Q, Qt, Qr -- same 9-tuple as the reference.  Vd, Z (DG1 spaces) are never
    used by the time loops and are returned as None; Qr is Q (same element, fenics_fem.py:286)."""
    if order != 2:
        raise NotImplementedError("the reference scripts run with order = 2")
    dm = dofmaps or {}
    S = {k: FunctionSpace(mesh, k, dm.get(k)) for k in ("W", "V", "Zd", "Zc", "Q", "Qt")}
    return S["W"], S["V"], None, None, S["Zd"], S["Zc"], S["Q"], S["Qt"], S["Q"]
