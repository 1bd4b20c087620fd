"""CPU: the mesh / dofmap ingestion path a dolfin user takes (fenics_b200.mesh.from_dolfin, function_spaces(mesh,
dofmaps=...)) on a duck-typed dolfin mesh: cells in any vertex order are brought to UFC order, exterior
facets and their local numbers follow, foreign dof numberings are taken as given.  (Against a real dolfin numbering:
tests/test_dolfin_harness.py, which needs dolfin.)"""
import numpy as np

from fenics_b200 import mesh as pmesh, spaces


class _DolfinMesh:
    def __init__(self, x, cells):
        self._x, self._c = x, cells

    def coordinates(self):
        return self._x

    def cells(self):
        return self._c


def test_from_dolfin_and_foreign_dofmaps():
    ref = pmesh.Skew_Mesh(5, 1, 1)
    rng = np.random.default_rng(0)
    cells = np.array([rng.permutation(c) for c in ref.cells()])          # vertex order within a cell scrambled
    m = pmesh.from_dolfin(_DolfinMesh(ref.coordinates().copy(), cells))
    assert np.array_equal(m.cells(), ref.cells()) and np.array_equal(m.coordinates(), ref.coordinates())
    assert np.array_equal(m.bfacet_cell, ref.bfacet_cell) and np.array_equal(m.bfacet_local, ref.bfacet_local)
    assert abs(m.hmin() - ref.hmin()) < 1e-15
    own = [s for s in spaces.function_spaces(m) if s is not None]
    perm = {s.name: rng.permutation(s.dim()) for s in own}
    given = {s.name: perm[s.name][s.cell_dofs].astype(np.int32) for s in own}
    for s in spaces.function_spaces(m, dofmaps=given):
        if s is None:
            continue
        assert np.array_equal(s.cell_dofs, given[s.name]) and s.dim() == len(perm[s.name])
        # dof coordinates follow the foreign numbering: dof perm[i] sits where the library's dof i sat
        mine = [t for t in own if t.name == s.name][0]
        assert np.allclose(s.tabulate_dof_coordinates()[perm[s.name]], mine.tabulate_dof_coordinates())
