"""CPU: how dolfin's preconditioner names resolve (fenics_b200/la.py:resolve_pc) -- `solve(A, x, b, "bicgstab", "default")`
(lid-driven-cavity/incomp_LDC.py:331), `KrylovSolver("cg", "amg")` (fenepmp_jbp.py:437-445).  No device needed."""
import warnings

import pytest

from fenics_b200 import la


class _Plan:
    def __init__(self, agg):
        self._agg = agg

    def has_aggregates(self, space):
        return self._agg


class _A:
    def __init__(self, agg=False):
        self.plan, self.space = _Plan(agg), "Q"


def test_names_resolve_loudly_and_default_can_mean_ilu():
    la._warned.clear()
    for name in ("none", "jacobi", "ilu", "jacobi_coarse"):
        assert la.resolve_pc(name, "bicgstab", _A()) == name              # implemented names are taken literally
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert la.resolve_pc("default", "bicgstab", _A()) == "jacobi"
        assert la.resolve_pc("default", "bicgstab", _A()) == "jacobi"     # one warning per name
        assert la.resolve_pc("amg", "cg", _A(agg=True)) == "jacobi_coarse"
        assert la.resolve_pc("hypre_amg", "bicgstab", _A(agg=True)) == "jacobi"   # the coarse correction is CG-only
    msgs = [str(x.message) for x in w if issubclass(x.category, la.PreconditionerSubstitution)]
    assert len(msgs) == 3 and "default_preconditioner" in msgs[0] and "no AMG kernel" in msgs[1]
    la.parameters["default_preconditioner"] = "ilu"                       # PETSc's serial default, literally
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            assert la.resolve_pc("default", "bicgstab", _A()) == "ilu"
    finally:
        la.parameters["default_preconditioner"] = "jacobi"
    la.parameters["preconditioner_substitution"] = "error"
    try:
        with pytest.raises(ValueError):
            la.resolve_pc("default", "bicgstab", _A())
    finally:
        la.parameters["preconditioner_substitution"] = "warn"
    with pytest.raises(ValueError):
        la.resolve_pc("sor", "bicgstab", _A())
    la._warned.clear()
