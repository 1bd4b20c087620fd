"""CPU: the literal-form registry closed numerically.  The loop-body TEXT of a reference script is executed against the
drop-in names; every form it hands to `assemble` / `project` is (i) recognised -> fx_form id, field slots, parameters
(fenics_b200/symbolic.py, form_registry.py) and assembled by that functor through the host emulation of the CUDA element
kernels (tests/hostemu: the same csrc headers), and (ii) translated node by node into the oracle's UFL restatement
(oracle/ufl_lite.py) and assembled there on the same random data.  Equal to 1e-12: the structure the registry keys on
really is the operator the hand-written functor computes -- for buoyancy-driven-flow/incompressible_dgp.py:369-523
(SURVEY.md 8f.2), whose literal path has no GPU comparison of its own."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import fxt_common as lc
import test_symbolic as ts
from fenics_b200 import _abi as abi
from fenics_b200 import form_registry, symbolic
from fxt_emu import EmuPlan
from oracle import fem
from oracle import ufl_lite as ufl
from oracle.configs_dgp import IncompDGP, dgp_mesh

TOL = 1e-12


class Translator:
    """symbolic.E tree -> oracle expression; template Functions -> oracle Functions with random data"""

    def __init__(self, spaces, seed=0):
        self.S, self.rng, self.fn = spaces, np.random.default_rng(seed), {}

    def function(self, tmpl, const=None):
        if id(tmpl) not in self.fn:
            sp_ = self.S[tmpl.function_space().name]
            f = fem.Function(sp_)
            f.vec[:] = self.rng.standard_normal(sp_.dim)
            if sp_ is self.S["Qt"]:
                f.vec[:] = np.abs(f.vec)                      # kappa, |res_orth|^2: non-negative by construction
            if const is not None:
                f.vec[:] = const
            self.fn[id(tmpl)] = (f, tmpl)
        return self.fn[id(tmpl)][0]

    def expr(self, e):
        op, X = e.op, self.expr
        a = e.a
        if op == "num":
            return ufl.as_expr(float(e.data))
        if op == "par":
            return ufl.as_expr(float(e.data.value))
        if op == "arg":
            number, space, comps, _ = e.data
            return ufl.Arg(self.S[space], number, comps, e.shape == ())
        if op == "coef":
            fn, comps = e.data
            return ufl.Coef(self.function(fn), comps, e.shape == ())
        if op == "h":
            return ufl.CellDiameter()
        if op == "n":
            return ufl.FacetNormal()
        if op == "zero":
            return ufl.Zero(e.shape)
        if op == "identity":
            return ufl.Identity(e.data)
        if op == "add":
            return X(a[0]) + X(a[1])
        if op == "sub":
            return X(a[0]) - X(a[1])
        if op == "mul":
            return X(a[0]) * X(a[1])
        if op == "div":
            return X(a[0]) / X(a[1])
        if op == "neg":
            return -X(a[0])
        if op == "pow":
            assert a[1].op in ("num", "par")
            return X(a[0]) ** float(a[1].data if a[1].op == "num" else a[1].data.value)
        if op == "sqrt":
            return ufl.sqrt(X(a[0]))
        if op == "abs":
            return ufl.ufl_abs(X(a[0]))
        if op == "dot":
            return ufl.dot(X(a[0]), X(a[1]))
        if op == "inner":
            return ufl.inner(X(a[0]), X(a[1]))
        if op == "grad":
            return ufl.grad(X(a[0]))
        if op == "nabla_grad":
            return ufl.nabla_grad(X(a[0]))
        if op == "divop":
            return ufl.div(X(a[0]))
        if op == "tr":
            return ufl.tr(X(a[0]))
        if op == "T":
            return ufl.transpose(X(a[0]))
        if op == "sym":
            m = X(a[0])
            return 0.5 * (m + ufl.transpose(m))
        if op == "idx":
            m = X(a[0])
            return m[e.data if len(e.data) > 1 else e.data[0]]
        if op == "vec":
            return ufl.as_vector([X(c) for c in a])
        if op == "mat":
            nc_ = e.data
            return ufl.as_matrix([[X(c) for c in a[r * nc_:(r + 1) * nc_]] for r in range(len(a) // nc_)])
        raise NotImplementedError(op)

    def form(self, form):
        out = None
        for e, kind, sub in form.integrals:
            m = ufl.Measure(kind, sub)
            term = self.expr(e) * m
            out = term if out is None else out + term
        return out


def test_incompressible_dgp_literal_forms_equal_their_functors_numerically():
    if not os.path.exists(ts.REF6):
        pytest.skip("reference checkout not present")
    mesh = dgp_mesh(5)
    cfg = IncompDGP(mesh, Ra=700.0, We=0.4, Pr=1.7, betav=0.3, th=0.8, Vh=0.02, Bi=0.3)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local, mesh.bfacet_marker)
    rec = []
    ns = ts._namespace6(rec)
    P = symbolic.Parameter
    ns.update(Ra=P("Ra", 700.0), Pr=P("Pr", 1.7), We=P("We", 0.4), dt=P("dt", 0.002), betav=P("betav", 0.3), th=P("th", 0.8),
              Vh=P("Vh", 0.02), Bi=P("Bi", 0.3), c1=P("c1", 0.07), c2=P("c2", 0.003), T_0=P("T0", 300.0), T_h=P("th_hot", 350.0))
    ns["thetal"] = ns["T"] / (ns["T_h"] - ns["T_0"])
    ns["theta0"] = (ns["T0"] - ns["T_0"]) / (ns["T_h"] - ns["T_0"])
    src = open(ts.REF6).read().splitlines()
    import textwrap
    exec(textwrap.dedent("\n".join(src[325:329])), ns)            # DEVSS terms again, with this test's parameters
    tr = Translator(cfg.S)
    tr.function(ns["thetar"], const=300.0 / (350.0 - 300.0))      # project(T_0/(T_h-T_0), Q): the constant at every dof
    for name in ("T0", "T1"):
        f = tr.function(ns[name])
        f.vec[:] = 325.0 + 10.0 * f.vec
    checked = []

    def both(form):
        fid, coefs, params, what, _ = symbolic.recognise(form)
        st = {slot: tr.function(t).vec for slot, t in coefs if slot >= 0}
        par = {getattr(abi, "FX_P_" + k.upper()): float(v.value) for k, v in params.items()}
        spaces = sorted(form.signature()[1].spaces)
        ref = fem.assemble(tr.form(form))
        return fid, st, par, spaces, ref, what

    def assemble(form):
        fid, st, par, spaces, ref, what = both(form)
        if sp.issparse(ref):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, spaces[0], st, par), ref)
        elif np.ndim(ref) == 1:
            err = lc.rel_err(plan.assemble_vector(fid, spaces[0], st, par), ref)
        else:
            got = plan.assemble_scalar(fid, st, par)
            err = abs(got - ref) / max(abs(ref), 1e-300)
        assert err < TOL, (what, err)
        checked.append(what)
        return ("tensor", fid)

    def project(expr, V):
        w = form_registry.arguments(V, 0)
        fid, st, par, spaces, ref, what = both(symbolic.inner(w, expr) * symbolic.dx)
        if isinstance(fid, tuple):                                  # DG0 target: cell averages, one kernel
            mass = fem.assemble(tr.form(symbolic.inner(w, form_registry.arguments(V, 1)) * symbolic.dx)).diagonal()
            err = lc.rel_err(plan.project_dg0(fid[1], V, st, par), ref / mass)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, V, st, par), ref)
        assert err < TOL, (what, err)
        checked.append(what)
        return ns["_projected"][fid]
    ns.update(assemble=assemble, project=project)
    exec("if True:\n" + "\n".join(src[368:523]) + "\n", ns)
    assert len(checked) == 20, checked
