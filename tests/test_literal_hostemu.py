"""CPU: the literal-form registry closed numerically.  The loop-body TEXT of a reference script is executed against the
drop-in names; every form it hands to `assemble` / `project` is (i) recognised -> fx_form id, field slots, parameters
(fenics_b200/symbolic.py, form_registry.py) and assembled by that functor through the host emulation of the CUDA element
kernels (tests/hostemu: the same csrc headers), and (ii) translated node by node into the oracle's UFL restatement
(oracle/ufl_lite.py) and assembled there on the same random data.  Equal to 1e-12, and the functor's quadrature degrees
equal UFL's estimate for the script's form: the structure the registry keys on really is the operator the hand-written
functor computes.  All eight scripts: configs 1-5, comp_LDC.py, comp_ewm_jpb.py, incompressible_dgp.py (the last two have
no GPU comparison of their literal path of their own).

Second half of the module: the COMPLETE loop bodies of all eight scripts, executed K times from their text against the
drop-in module's names with `assemble` / `project` served by the recognised functors (host emulation), `bc.apply` and
exact LU on the host, follow the oracle's own step() trajectory -- the drop-in claim without a GPU."""
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
_HIT = set()      # registry signatures the tests of this module have closed numerically (see the last test)


class Translator:
    """symbolic.E tree -> oracle expression; template Functions -> oracle Functions with random data"""

    def __init__(self, spaces, seed=0, by_slot=None):
        self.S, self.rng, self.fn, self.by_slot = spaces, np.random.default_rng(seed), {}, dict(by_slot or {})

    def function(self, tmpl, const=None):
        """the oracle Function standing behind a template: the randomised field of the oracle config in that slot (value
        ranges chosen by the config's own test helpers), else random data"""
        if id(tmpl) not in self.fn:
            f = self.by_slot.get(getattr(tmpl, "slot", None))
            if f is None:
                sp_ = self.S[tmpl.function_space().name]
                f = fem.Function(sp_)
                f.vec[:] = self.rng.standard_normal(sp_.dim)
                if sp_ is self.S["Qt"]:
                    f.vec[:] = np.abs(f.vec)                  # kappa, |res_orth|^2: non-negative by construction
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



def _close(ns, tr, plan, projected=None, base_par=None):
    """install assemble / project that check every form both ways; returns the list of statements checked.  base_par: the
    parameter table as the run's earlier statements left it (a Problem keeps its parameters between forms: a functor
    that folds a projected constant -- thetar = T_0/(T_h-T_0) -- reads T_0, T_h although the form no longer names them)"""
    checked = []

    def both(form):
        fid, coefs, params, what, _ = symbolic.recognise(form)
        _HIT.add(form.signature()[0])
        st = {slot: tr.function(t).vec for slot, t in coefs if slot >= 0}
        par = dict(base_par or {})
        par.update({getattr(abi, "FX_P_" + k.upper()): float(v.value) for k, v in params.items()})
        spaces = sorted(form.signature()[1].spaces)
        oform = tr.form(form)
        ref = fem.assemble(oform)
        if not isinstance(fid, tuple):          # ... integrated with the rule UFL's degree estimation gives the script's form
            degs = {}
            for k, _, d in fem.quadrature_degrees(oform, plan.mesh):
                degs[k] = max(degs.get(k, -1), d)
            q = plan.form_qdeg(fid, par.get(abi.FX_P_CONV, 1.0))
            assert q[0] == degs.get("dx", -1) or "dx" not in degs, (what, q, degs)
            assert q[1] == degs.get("ds", -1) or "ds" not in degs, (what, q, degs)
        return fid, st, par, spaces, ref, what

    def assemble(form):
        fid, st, par, spaces, ref, what = both(form)
        if sp.issparse(ref):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, spaces[0], st, par), ref)
        elif np.ndim(ref) == 1:
            err = lc.rel_err(plan.assemble_vector(fid, spaces[0], st, par), ref)
        else:
            err = abs(plan.assemble_scalar(fid, st, par) - ref) / max(abs(ref), 1e-300)
        assert err < TOL, (what, err)
        checked.append(what)
        return ("tensor", fid)

    def project(expr, V):
        w = form_registry.arguments(V, 0)
        try:
            fid, st, par, spaces, ref, what = both(symbolic.inner(w, expr) * symbolic.dx)
        except symbolic.UnknownForm:                                # compressible_dgp.py:404, dead in the script
            return None
        if isinstance(fid, tuple):                                  # DG0 target: cell averages, one kernel
            mass = fem.assemble(tr.form(symbolic.inner(w, form_registry.arguments(V, 1)) * symbolic.dx)).diagonal()
            err = lc.rel_err(plan.project_dg0(fid[1], V, st, par), ref / mass)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, V, st, par), ref)
            # ... next to the consistent mass matrix dolfin.project assembles (dolfin_api.project: recognised like any form)
            mid, _, mpar, _, mref, mwhat = both(symbolic.inner(w, form_registry.arguments(V, 1)) * symbolic.dx)
            assert lc.csr_rel_err(plan.assemble_matrix(mid, V, {}, mpar), mref) < TOL, mwhat
        assert err < TOL, (what, err)
        checked.append(what)
        return projected[fid] if projected is not None and fid in projected else tr_result(V, fid)

    def tr_result(V, fid):
        raise KeyError("project(): no template registered for the result of %r on %s" % (fid, V))
    ns.update(assemble=assemble, project=project)
    return checked


def _lps_results(restau_expr, res_orth_form):
    """templates standing for the results of the four project() calls of an LPS indicator block"""
    A, T = abi, ts.T
    return {("expr", restau_expr): T("Zd", A.FX_F_RES_TEST), res_orth_form: T("Zc", A.FX_F_RES_ORTH),
            ("expr", A.FX_EXPR_RES_ORTH_SQ): T("Qt", A.FX_F_QT_A), ("expr", A.FX_EXPR_SQRT_QT_A): T("Qt", A.FX_F_KAPP)}


def _params_from(ns, values):
    """the namespace's Parameters take the oracle config's values (same objects: forms built earlier see them)"""
    for v in ns.values():
        if isinstance(v, symbolic.Parameter) and v.name in values:
            v.value = float(values[v.name])


def _plan(mesh, cfg):
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local, mesh.bfacet_marker)
    plan.mesh = mesh
    return plan


def test_incompressible_dgp_literal_forms_equal_their_functors_numerically():
    """buoyancy-driven-flow/incompressible_dgp.py:369-523"""
    if not os.path.exists(ts.REF6):
        pytest.skip("reference checkout not present")
    import fxt_dgp as dg
    mesh = dgp_mesh(5)
    cfg = IncompDGP(mesh, Ra=700.0, We=0.4, Pr=1.7, betav=0.3, th=0.8, Vh=0.02, Bi=0.3, dt=0.002, c1=0.07, c2=0.003)
    dg.c6_randomize(cfg)
    ns = ts._namespace6([])
    _params_from(ns, cfg.p)
    tr = Translator(cfg.S, by_slot={slot: getattr(cfg, a) for slot, a in dg.C6_FIELD_ATTR.items()})
    tr.function(ns["thetar"], const=cfg.p["T0"] / (cfg.p["th_hot"] - cfg.p["T0"]))   # project(T_0/(T_h-T_0), Q) :286-287
    checked = _close(ns, tr, _plan(mesh, cfg), ns["_projected"],
                     base_par={abi.FX_P_T0: cfg.p["T0"], abi.FX_P_TH_HOT: cfg.p["th_hot"]})
    src = open(ts.REF6).read().splitlines()
    exec("if True:\n" + "\n".join(src[368:523]) + "\n", ns)
    assert len(checked) == 20, checked


def test_config1_literal_forms_equal_their_functors_numerically():
    """lid-driven-cavity/incomp_LDC.py:298-308 (LPS block), :317-421 (five systems, energies)"""
    if not os.path.exists(ts.REF):
        pytest.skip("reference checkout not present")
    from oracle import configs
    mesh = fem.skew_mesh(5)
    cfg = configs.IncompLDC(mesh)
    lc.randomize(cfg)
    ns = ts._namespace([])
    _params_from(ns, cfg.p)
    T = ts.T
    tr = Translator(cfg.S, by_slot={slot: getattr(cfg, a) for slot, a in lc.FIELD_ATTR.items() if hasattr(cfg, a)})
    projected = {("expr", abi.FX_EXPR_LDC_RESTAU): T("Zd", abi.FX_F_RES_TEST), abi.FX_FORM_LDC_L_RES_ORTH: T("Zc", abi.FX_F_RES_ORTH),
                 ("expr", abi.FX_EXPR_RES_ORTH_SQ): T("Qt", abi.FX_F_QT_A), ("expr", abi.FX_EXPR_SQRT_QT_A): T("Qt", abi.FX_F_KAPP)}
    ns.update(Zd="Zd", Zc="Zc", Qt="Qt", np=np)
    checked = _close(ns, tr, _plan(mesh, cfg), projected)
    projected[("expr", abi.FX_EXPR_H_KAPP)] = T("Qt", -1)
    for a, b in ((298, 308), (317, 331), (355, 363), (370, 375), (379, 394), (400, 416), (420, 421), (434, 434)):
        exec(ts._lines(a, b), ns)
    assert len(checked) == 4 + 10 + 2 + 1, checked


def test_config2_literal_forms_equal_their_functors_numerically():
    """lid-driven-cavity/comp_LDC_mesh_convergence.py:417-549 (six systems through lhs / rhs, the density projection,
    energies) and the two momentum steps of the sweep script comp_LDC.py (Re*conv -> the functor's conv slot)"""
    if not os.path.exists(ts.REF2):
        pytest.skip("reference checkout not present")
    from oracle import configs
    mesh = fem.skew_mesh(5)
    cfg = configs.CompLDC(mesh)
    lc.randomize(cfg)
    ns = ts._namespace2([])
    _params_from(ns, cfg.p)
    by_slot = {slot: getattr(cfg, a) for slot, a in lc.FIELD_ATTR.items() if hasattr(cfg, a)}
    tr = Translator(cfg.S, by_slot=by_slot)
    projected = {abi.FX_FORM_C2_L_RHO1: ts.T("Q", abi.FX_F_RHO1)}
    ns.update(Q="Q")
    checked = _close(ns, tr, _plan(mesh, cfg), projected)
    for a, b in ((417, 426), (430, 447), (458, 476), (483, 497), (501, 502), (506, 519), (528, 544), (548, 549)):
        exec(ts._lines2(a, b), ns)
    assert len(checked) == 12 + 1 + 2, checked
    ref3 = os.path.join(os.path.dirname(ts.REF), "comp_LDC.py")
    src = open(ref3).read().splitlines()
    import textwrap
    for a, b in ((418, 441), (468, 489)):
        ns = ts._namespace2([])
        _params_from(ns, dict(cfg.p, Re=25.0, conv=0.5))
        checked = _close(ns, Translator(cfg.S, by_slot=by_slot), _plan(mesh, cfg))
        exec(textwrap.dedent("\n".join(src[a - 1:b])), ns)
        assert len(checked) == 2, checked


def test_config5_literal_forms_equal_their_functors_numerically():
    """buoyancy-driven-flow/compressible_dgp.py:404-583"""
    if not os.path.exists(ts.REF5):
        pytest.skip("reference checkout not present")
    import fxt_dgp as dg
    from oracle.configs_dgp import CompDGP
    mesh = dgp_mesh(5)
    cfg = CompDGP(mesh)
    dg.randomize(cfg)
    ns = ts._namespace5([])
    _params_from(ns, cfg.p)
    A = abi
    tr = Translator(cfg.S, by_slot={slot: getattr(cfg, a) for slot, a in dg.FIELD_ATTR.items()})
    tr.function(ns["thetar"], const=cfg.p["T0"] / (cfg.p["th_hot"] - cfg.p["T0"]))
    projected = {A.FX_FORM_C5_L_THETA0: ts.T("Q", A.FX_F_THETA0), A.FX_FORM_C5_L_THERM_INV: ts.T("Q", A.FX_F_THERM_INV),
                 A.FX_FORM_C5_L_RHO1: ts.T("Q", A.FX_F_RHO1)}
    calls = {"n": 0}
    first = projected[A.FX_FORM_C5_L_THETA0]

    class _Proj(dict):                     # :581 projects theta1 with theta0's functor: its result feeds the Nusselt integral
        def __getitem__(self, fid):
            if fid == A.FX_FORM_C5_L_THETA0:
                calls["n"] += 1
                return first if calls["n"] == 1 else ts.T("Q", A.FX_F_Q_A)
            return dict.__getitem__(self, fid)
    projected.update(_lps_results(A.FX_EXPR_C5_RESTAU, A.FX_FORM_C5_L_RES_ORTH))
    ns.update(Zd="Zd", Zc="Zc", Qt="Qt", np=np, I_vec=symbolic.as_vector([1.0, 0.0, 1.0]))
    checked = _close(ns, tr, _plan(mesh, cfg), _Proj(projected),
                     base_par={A.FX_P_T0: cfg.p["T0"], A.FX_P_TH_HOT: cfg.p["th_hot"]})
    src = open(ts.REF5).read().splitlines()
    exec("if True:\n" + "\n".join(src[381:391]) + "\n", ns)        # the LPS indicator block :382-391
    assert len(checked) == 4, checked
    ns["kapp"] = ts.T("Qt", A.FX_F_KAPP)
    exec("if True:\n" + "\n".join(src[403:583]) + "\n", ns)
    assert len(checked) == 4 + 16 + 4 + 3, checked


@pytest.mark.parametrize("mp", [True, False])
def test_config4_literal_forms_equal_their_functors_numerically(mp):
    """"flow between eccentrically rotating cylinders"/fenepmp_jbp.py:489-658, FENE-P-MP (lambda_d = 0.1) and FENE-P"""
    if not os.path.exists(ts.REF4):
        pytest.skip("reference checkout not present")
    import fxt_jbp as jb
    from oracle.configs_jbp import FenepmpJBP
    mesh = fem.annulus_mesh(12, 3)
    cfg = FenepmpJBP(mesh, lambda_d=0.1 if mp else 0.0)
    jb.randomize(cfg)
    ns = ts._namespace4([], symbolic.Parameter("lambda_d", 0.1) if mp else 0.0)
    _params_from(ns, cfg.p)
    tr = Translator(cfg.S, by_slot={slot: getattr(cfg, a) for slot, a in jb.FIELD_ATTR.items()})
    tr.function(ns["thetar"], const=cfg.p["T0"] / (cfg.p["th_hot"] - cfg.p["T0"]))
    A = abi
    projected = _lps_results(A.FX_EXPR_C4M_RESTAU if mp else A.FX_EXPR_C4_RESTAU,
                             A.FX_FORM_C4M_L_RES_ORTH if mp else A.FX_FORM_C4_L_RES_ORTH)
    ns.update(Zd="Zd", Zc="Zc", Qt="Qt", np=np, I_vec=symbolic.as_vector([1.0, 0.0, 1.0]), absolute=lambda k: k)
    checked = _close(ns, tr, _plan(mesh, cfg), projected,
                     base_par={abi.FX_P_T0: cfg.p["T0"], abi.FX_P_TH_HOT: cfg.p["th_hot"]})
    src = open(ts.REF4).read().splitlines()
    for a, b in ((406, 407), (413, 413), (415, 415)):
        exec("if True:\n" + "\n".join(src[a - 1:b]) + "\n", ns)
    exec("if True:\n" + "\n".join(src[460:473]) + "\n", ns)        # the LPS indicator block :461-473
    assert len(checked) == 4, checked
    exec("if True:\n" + "\n".join(src[488:658]) + "\n", ns)
    assert len(checked) == 4 + 19, checked


def _run_ewm(ref, cfg, mesh, ns, setup_ranges, body, n_expected):
    import fxt_jbp as jb
    _params_from(ns, cfg.p)
    tr = Translator(cfg.S, by_slot={slot: getattr(cfg, a) for slot, a in jb.FIELD_ATTR.items() if hasattr(cfg, a)})
    tr.function(ns["thetar"], const=cfg.p["T0"] / (cfg.p["th_hot"] - cfg.p["T0"]))
    checked = _close(ns, tr, _plan(mesh, cfg), base_par={abi.FX_P_T0: cfg.p["T0"], abi.FX_P_TH_HOT: cfg.p["th_hot"]})
    src = open(ref).read().splitlines()
    import textwrap
    for a, b in setup_ranges:
        exec(textwrap.dedent("\n".join(src[a - 1:b])), ns)
    exec("if True:\n" + "\n".join(src[body[0]:body[1]]) + "\n", ns)
    for name in ("innertorque", "innerforcex", "innerforcey"):      # assembled behind the scripts' record blocks
        ns["assemble"](ns[name])
    assert len(checked) == n_expected, checked


@pytest.mark.parametrize("general", [True, False])
def test_config3_literal_forms_equal_their_functors_numerically(general):
    """incompressible_benchmarkEWM.py:488-706 with the material functions as Parameters (C3E functors) and with the script's
    plain 0.0 (folded C3 functors)"""
    if not os.path.exists(ts.REF3):
        pytest.skip("reference checkout not present")
    import fxt_ewm as ew
    from oracle.configs_ewm import EwmJBP
    mesh = fem.annulus_mesh(12, 3, x_1=-0.9)
    kw = dict(ew.GENERAL, A_0=0.4, k_ewm=0.7, B=0.2) if general else dict(ew.GENERAL)
    cfg = EwmJBP(mesh, **kw)
    ew.randomize(cfg)
    ns = ts._namespace4([], 0.0)
    P = symbolic.Parameter
    ns.update(al=P("al1", 0.001), adaptive_loop=False, err_count=0, max_refine=2, primary_loop=1, np=np,
              A_0=P("a0", 0.4) if general else 0.0, k_ewm=P("k_ewm", 0.7) if general else 0.0,
              B=P("b_ewm", 0.2) if general else 0.0, norm=lambda *a: 0.0)

    class _V:
        def get_local(self):
            return np.zeros(1)
    for f in ("w0", "w1", "tau0_vec", "tau1_vec"):
        ns[f].vector = lambda: _V()
    _run_ewm(ts.REF3, cfg, mesh, ns, ((402, 402), (410, 411), (415, 416), (422, 422), (472, 475)), (487, 706), 18 + 3)


@pytest.mark.parametrize("general", [True, False])
def test_comp_ewm_jpb_literal_forms_equal_their_functors_numerically(general):
    """comp_ewm_jpb.py:426-659: the true extended White-Metzner loop (and with A_0 = k_ewm = B = 0.0, where
    the material functions fold away)"""
    ref = os.path.join(os.path.dirname(ts.REF3), "comp_ewm_jpb.py")
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    import fxt_ewm as ew
    from oracle.configs_ewm import CompEwmJBP
    mesh = fem.annulus_mesh(12, 3, x_1=-0.8)
    cfg = CompEwmJBP(mesh, **{k: v for k, v in ew.EWM_GENERAL.items()})
    ew.randomize(cfg)
    ns = ts._namespace4([], 0.0)
    P = symbolic.Parameter
    ns.update(al=P("al1", 0.001), np=np, norm=lambda *a: 0.0, iter=1,
              A_0=P("a0", 120.0) if general else 0.0, k_ewm=P("k_ewm", -0.7) if general else 0.0,
              B=P("b_ewm", 0.1) if general else 0.0)

    class _V:
        def get_local(self):
            return np.zeros(1)
    for f in ("w0", "w1", "tau0_vec", "tau1_vec"):
        ns[f].vector = lambda: _V()
    _run_ewm(ref, cfg, mesh, ns, ((368, 368), (376, 377), (381, 382), (386, 389)), (425, 659), 18 + 3)


def test_every_registered_signature_is_closed_numerically():
    """nothing in the registry is there on trust: run after the tests above (whole module), every signature of
    form_registry.py has been assembled both ways"""
    form_registry.load()
    if len(_HIT) < 60:
        pytest.skip("run the whole module: this test sums up the others")
    missing = [v[2] for k, v in symbolic._REGISTRY.items() if k not in _HIT]
    assert not missing, missing


# ---- the whole loop body, literally, on the functors -----------------------------------------------------------------------
class _EmuVector:
    def __init__(self, fn):
        self.fn = fn

    def get_local(self):
        return self.fn.vec


class _EmuFunction(symbolic.Operand):
    """CPU stand-in for dolfin_api.Function: a symbolic operand backed by an oracle Function's vector"""

    def __init__(self, fn, name):
        self.fn, self._space = fn, form_registry._Space(name)

    def function_space(self):
        return self._space

    def split(self):
        return form_registry.split(self)

    def vector(self):
        return _EmuVector(self.fn)

    def assign(self, other):
        self.fn.vec[:] = other.fn.vec


def _number(e):
    """`t += dt` with dt a Parameter leaves a constant tree: its value (symbolic.E.__float__)"""
    return float(e)


class _Clock:                           # the progress bar's datetime.timedelta(seconds=...) gets such trees too
    @staticmethod
    def timedelta(seconds=0):
        return _number(seconds)


class _Setter:
    """Expression with a settable .t (ulidreg.t = t): forwards it to the oracle's boundary-value closure"""

    def __init__(self, holder=None, key="t"):
        object.__setattr__(self, "_h", (holder, key))

    def __setattr__(self, k, v):
        holder, key = self._h
        if k == "t" and holder is not None:
            holder[key] = _number(v)

    def __getattr__(self, k):          # `str(ramp_function.t)` in compressible_dgp.py's progress line
        holder, key = object.__getattribute__(self, "_h")
        if k == "t":
            return holder[key] if holder is not None else 0.0
        raise AttributeError(k)


class _EmuSolver:
    def __init__(self, solve):
        self._solve = solve

    def solve(self, A, x, b):
        self._solve(A, x, b)


def _emu_runtime(ff, cfg, plan, fields, params):
    """namespace of a script's set-up section on the CPU stand-ins: the drop-in module's names, Parameters, Functions backed
    by the oracle config's vectors, and assemble / project / solve served by the recognised functors + host LU"""
    import time
    ns = {k: getattr(ff, k) for k in dir(ff) if not k.startswith("_")}
    ns.update({script_name: symbolic.Parameter(slot_name, cfg.p[slot_name]) for script_name, slot_name in params.items()})
    for name, space in fields.items():
        ns[name] = _EmuFunction(getattr(cfg, name), space)
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, ns["p"], _, tau_vec, ns["u"], D_vec, ns["D"], ns["tau"] = ff.trial_functions("Q", "Zc", "W")
    ns["rho"] = ns["T"] = ns["p"]
    ns["q"] = ns["r"] = ff.TestFunction("Q")
    (ns["v"], R_vec), Rt_vec = ff.TestFunctions("W"), ff.TestFunction("Zc")
    ns["R"], ns["Rt"] = m3(R_vec), m3(Rt_vec)
    (ns["u0"], D0_vec), (ns["u12"], D12_vec), (ns["us"], _), (ns["u1"], _) = (ns[k].split() for k in ("w0", "w12", "ws", "w1"))
    ns.update(D0=m3(D0_vec), D12=m3(D12_vec), tau0=m3(symbolic.as_expr(ns["tau0_vec"])), tau1=m3(symbolic.as_expr(ns["tau1_vec"])),
              h=symbolic.CellDiameter(), n=symbolic.FacetNormal(), Q="Q", Zd="Zd", Zc="Zc", Qt="Qt", np=np,
              I_vec=symbolic.as_vector([1.0, 0.0, 1.0]))
    par = {}

    def run(form):
        fid, coefs, params_, what, _ = symbolic.recognise(form)
        par.update({getattr(abi, "FX_P_" + k.upper()): float(v.value) for k, v in params_.items()})   # a Problem keeps them
        st = {slot: f.fn.vec for slot, f in coefs if slot >= 0}
        return fid, st, sorted(form.signature()[1].spaces), len(form.signature()[1].args)

    def assemble(form):
        fid, st, spaces, rank = run(form)
        if rank == 2:
            return plan.assemble_matrix(fid, spaces[0], st, par)
        if rank == 1:
            return plan.assemble_vector(fid, spaces[0], st, par)
        return plan.assemble_scalar(fid, st, par)

    def project(expr, V):
        w = form_registry.arguments(V, 0)
        try:
            fid, st, _, _ = run(symbolic.inner(w, expr) * symbolic.dx)
        except symbolic.UnknownForm:                                 # compressible_dgp.py:404: dead in the script
            return None
        out = _EmuFunction(fem.Function(cfg.S[V]), V)
        if isinstance(fid, tuple):
            out.fn.vec[:] = plan.project_dg0(fid[1], V, st, par)
        else:
            M = assemble(symbolic.inner(w, form_registry.arguments(V, 1)) * symbolic.dx)
            out.fn.vec[:] = fem.solve_lu(M, plan.assemble_vector(fid, V, st, par))
        return out

    def solve(A, x, b, *a):
        x.fn.vec[:] = fem.solve_lu(A, b)

    def absolute(f):
        f.fn.vec[:] = np.abs(f.fn.vec)
        return f
    ns.update(assemble=assemble, project=project, solve=solve, absolute=absolute, solvertau=_EmuSolver(solve), end=lambda: None,
              bcu=cfg.bcu, bcp=[], bctau=[], time=time, datetime=_Clock, update_progress=lambda *a: None, jjj=0, j=0, Tf=10.0,
              elapsed=0.0, total_elapsed=0.0, iter=0, t=cfg.t, t_array=[], ek_array=[], ee_array=[], nus_array=[],
              ulid=_Setter(), rampd=_Setter(), Ret=_Setter(), Wet=_Setter())
    return ns


@pytest.mark.parametrize("script,cls,devss,rng", [("comp_LDC_mesh_convergence.py", "CompLDC", (343, 349), (377, 565)),
                                                   ("comp_LDC.py", "CompLDCSweep", (334, 340), (365, 594))])
def test_config2_loop_body_text_runs_on_the_functors_for_k_steps(script, cls, devss, rng):
    """comp_LDC_mesh_convergence.py:377-565 (and the sweep script comp_LDC.py:365-594, whose convective terms carry Re*conv)
    -- the bench configuration's COMPLETE loop body (LPS block with its four
    projections, six linear systems, density projection, energies, field roll, t += dt), written against the drop-in
    names, executed K times with `assemble` = recognise + the recognised functor (host emulation of the CUDA element
    kernels), `project` likewise, `bc.apply` / LU on the host: the trajectory of the oracle's own step().  (On the GPU the
    same text runs through dolfin_api: tests/test_gpu_literal.py.)"""
    if not os.path.exists(ts.REF2):
        pytest.skip("reference checkout not present")
    import textwrap
    import fenics_b200.shims.lid_driven_cavity.fenics_fem as ff
    from oracle import configs
    K = 3
    mesh = fem.skew_mesh(6)
    cfg, twin = getattr(configs, cls)(mesh), getattr(configs, cls)(mesh)
    cfg.t = twin.t = 0.45
    fields = dict(w0="W", w12="W", ws="W", w1="W", p0="Q", p1="Q", rho0="Q", rho12="Q", rho1="Q", tau0_vec="Zc", tau1_vec="Zc")
    ns = _emu_runtime(ff, cfg, _plan(mesh, cfg), fields, {k: k for k in cfg.p})
    ns["T0"], ns["T1"] = (_EmuFunction(fem.Function(cfg.S["Q"]), "Q") for _ in range(2))   # rolled by the script, unused (:560)
    ns.update(ulidreg=_Setter(cfg._bct), ramp_function=_Setter(), Tf=1.0)
    src = open(os.path.join(os.path.dirname(ts.REF2), script)).read().splitlines()
    exec(textwrap.dedent("\n".join(src[devss[0] - 1:devss[1]])), ns)   # DEVSS definitions before the loop
    body = "if True:\n" + "\n".join(src[rng[0] - 1:rng[1]]) + "\n"
    for _ in range(K):
        exec(body, ns)
        ns["t"] = _number(ns["t"])
        r = twin.step()
        assert abs(ns["E_k"] - r["E_k"]) <= 1e-10 * abs(r["E_k"]) and abs(ns["E_e"] - r["E_e"]) <= 1e-10 * max(abs(r["E_e"]), 1e-14)
    assert abs(ns["t"] - twin.t) < 1e-14 and ns["iter"] == K
    for name, ref in (("w1", twin.w1), ("p1", twin.p1), ("tau1_vec", twin.tau1_vec), ("w0", twin.w0), ("rho0", twin.rho0),
                      ("tau0_vec", twin.tau0_vec), ("p0", twin.p0), ("w12", twin.w12), ("ws", twin.ws)):
        assert lc.rel_err(ns[name].fn.vec, ref.vec) < 1e-9, name
    assert lc.rel_err(ns["rho1"].fn.vec, twin.rho1.vec) < 1e-9
    assert lc.rel_err(ns["kapp"].fn.vec, twin.kapp.vec) < 1e-7
    assert np.abs(twin.w1.vec).max() > 1e-3


def test_config5_loop_body_text_runs_on_the_functors_for_k_steps():
    """buoyancy-driven-flow/compressible_dgp.py:369-595 -- LPS block, three per-step projections, eight linear systems,
    density projection, energies, Nusselt number, the hot wall ramped in time: K steps from the text on the functors"""
    if not os.path.exists(ts.REF5):
        pytest.skip("reference checkout not present")
    import textwrap
    import fenics_b200.shims.buoyancy_driven_flow.fenics_fem as ff
    from oracle.configs_dgp import CompDGP
    K = 3
    mesh = dgp_mesh(6)
    cfg, twin = CompDGP(mesh), CompDGP(mesh)
    cfg.t = twin.t = 1.4                        # the hot wall is being ramped (tanh(4 (t - 1.5)))
    fields = dict(w0="W", w12="W", ws="W", w1="W", p0="Q", p1="Q", rho0="Q", rho12="Q", rho1="Q", tau0_vec="Zc", tau1_vec="Zc",
                  T0="Q", T1="Q", thetar="Q")
    names = dict(Ra="Ra", Pr="Pr", We="We", Ma="Ma", dt="dt", betav="betav", th="th", Vh="Vh", Bi="Bi", c1="c1", c2="c2",
                 T_0="T0", T_h="th_hot", al_1="al1", al_2="al2")
    ns = _emu_runtime(ff, cfg, _plan(mesh, cfg), fields, names)
    ns["thetar"].slot = -1                      # project(T_0/(T_h-T_0), Q) :295-296: a constant inside the functors
    ns.update(C=200.0, k=symbolic.as_vector([0.0, 1.0]), bcT=cfg.bcT, ramp_function=_Setter(cfg._tt))
    ns["thetal"] = ns["T"] / (ns["T_h"] - ns["T_0"])                  # :294
    src = open(ts.REF5).read().splitlines()
    exec(textwrap.dedent("\n".join(src[335:338])), ns)               # DEVSS definitions :336-338
    body = "if True:\n" + "\n".join(src[368:595]) + "\n"
    for _ in range(K):
        exec(body, ns)
        ns["t"] = _number(ns["t"])
        r = twin.step()
        for key in ("E_k", "E_e", "Nus"):
            assert abs(ns[key] - r[key]) <= 1e-9 * max(abs(r[key]), 1e-12), key
    assert abs(ns["t"] - twin.t) < 1e-14
    for name, ref in (("w1", twin.w1), ("p1", twin.p1), ("tau1_vec", twin.tau1_vec), ("T1", twin.T1), ("w0", twin.w0),
                      ("rho0", twin.rho0), ("T0", twin.T0), ("w12", twin.w12), ("ws", twin.ws), ("rho12", twin.rho12)):
        assert lc.rel_err(ns[name].fn.vec, ref.vec) < 1e-9, name
    assert lc.rel_err(ns["rho1"].fn.vec, twin.rho1.vec) < 1e-9
    assert lc.rel_err(ns["kapp"].fn.vec, twin.kapp.vec) < 1e-7
    assert np.abs(twin.w1.vec).max() > 1e-6


def test_config1_loop_body_text_runs_on_the_functors_for_k_steps():
    """lid-driven-cavity/incomp_LDC.py:281-447 -- LPS block, five linear systems incl. the singular pure-Neumann pressure
    system, energies, the error indicator and the divergence test with its `break`, field roll: K steps from the text on
    the functors"""
    if not os.path.exists(ts.REF):
        pytest.skip("reference checkout not present")
    import textwrap
    import fenics_b200.shims.lid_driven_cavity.fenics_fem as ff
    from oracle import configs
    K = 5
    mesh = fem.skew_mesh(6)
    cfg, twin = configs.IncompLDC(mesh), configs.IncompLDC(mesh)
    cfg.t = twin.t = 0.49                       # crosses t = 0.5, where the divergence test (:437-441) switches on
    fields = dict(w0="W", w12="W", ws="W", w1="W", p0="Q", p1="Q", tau0_vec="Zc", tau1_vec="Zc")
    ns = _emu_runtime(ff, cfg, _plan(mesh, cfg), fields, {k: k for k in cfg.p})

    def solve(A, x, b, *a):
        singular = abs(A @ np.ones(A.shape[0])).max() < 1e-12 * abs(A).max()
        x.fn.vec[:] = configs.solve_zero_sum(A, b) if singular else fem.solve_lu(A, b)   # bcp = []: pure Neumann (:227)
    ns["T0"], ns["T1"] = (_EmuFunction(fem.Function(cfg.S["Q"]), "Q") for _ in range(2))
    ns.update(solve=solve, norm=lambda v, kind: fem.norm(v.get_local(), kind), ulidreg=_Setter(cfg._bct), Tf=1.0, err_array=[])
    src = open(ts.REF).read().splitlines()
    exec(textwrap.dedent("\n".join(src[258:263])), ns)                # DEVSS / LPS definitions :259-263
    ns.update(a1=0, L1=0)
    raw = textwrap.dedent("\n".join(src[280:447]))
    body = "for once_ in (0,):\n" + textwrap.indent(raw, "    ") + "\n    completed_ = True\n"
    for step in range(K):
        ns["completed_"] = False
        exec(body, ns)
        ns["t"] = _number(ns["t"])
        r = twin.step()
        assert ns["completed_"] == (not r["diverged"])
        assert abs(ns["E_k"] - r["E_k"]) <= 1e-9 * abs(r["E_k"]) and abs(ns["E_e"] - r["E_e"]) <= 1e-9 * max(abs(r["E_e"]), 1e-14)
        assert abs(ns["err_array"][-1] - r["err"]) <= 1e-7 * max(r["err"], 1e-14)
        if r["diverged"]:
            break
    assert ns["t"] > 0.5 and abs(ns["t"] - twin.t) < 1e-14 and ns["iter"] == step + 1
    rolled = () if r["diverged"] else (("w0", twin.w0), ("tau0_vec", twin.tau0_vec))
    for name, ref in (("w1", twin.w1), ("tau1_vec", twin.tau1_vec), ("w12", twin.w12), ("ws", twin.ws)) + rolled:
        assert lc.rel_err(ns[name].fn.vec, ref.vec) < 1e-9, name
    pm = lambda v: v - v.mean()  # noqa: E731  (the pressure is defined up to a constant)
    assert lc.rel_err(pm(ns["p1"].fn.vec), pm(twin.p1.vec)) < 1e-8
    assert lc.rel_err(ns["kapp"].fn.vec, twin.kapp.vec) < 1e-7


_JBP_FIELDS = dict(w0="W", w12="W", ws="W", w1="W", p0="Q", p1="Q", rho0="Q", rho12="Q", rho1="Q", tau0_vec="Zc", tau1_vec="Zc",
                   T0="Q", T1="Q", thetar="Q")
_JBP_NAMES = dict(Re="Re", We="We", Ma="Ma", dt="dt", betav="betav", th="th", conv="conv", Di="Di", Vh="Vh", Bi="Bi", T_0="T0",
                  T_h="th_hot")


def _jbp_runtime(cfg, mesh, extra_names):
    import fenics_b200.shims.eccentric_cylinders.fenics_base as fb
    ns = _emu_runtime(fb, cfg, _plan(mesh, cfg), _JBP_FIELDS, dict(_JBP_NAMES, **extra_names))
    ns["thetar"].slot = -1
    ns.update(bcT=cfg.bcT, w=_Setter(cfg._tt), tang=symbolic.as_vector([ns["n"][1], -ns["n"][0]]), mesh_refinement=False,
              j=1, primary_loop=1, secondary_loop=0, x1=[], ek1=[], ee1=[], y1=[], zx1=[], z1=[],
              norm=lambda v, kind: fem.norm(v.get_local(), kind))
    ns["thetal"] = ns["T"] / (ns["T_h"] - ns["T_0"])
    ns["theta0"] = (ns["T0"] - ns["T_0"]) / (ns["T_h"] - ns["T_0"])
    return ns


def _jbp_compare(ns, twin, K, fields, keys):
    assert abs(ns["t"] - twin.t) < 1e-14 and ns["iter"] == K
    for name in fields:
        assert lc.rel_err(ns[name].fn.vec, getattr(twin, name).vec) < 1e-9, name
    assert np.abs(twin.w1.vec).max() > 1e-3


def test_config4_loop_body_text_runs_on_the_functors_for_k_steps():
    """fenepmp_jbp.py:455-690, FENE-P-MP (lambda_D = 0.1): LPS block with absolute(kapp), `if iter > 1` roll, seven linear
    systems, energies, journal torque / load -- K steps from the text on the functors"""
    if not os.path.exists(ts.REF4):
        pytest.skip("reference checkout not present")
    import textwrap
    from oracle.configs_jbp import FenepmpJBP
    K = 3
    mesh = fem.annulus_mesh(12, 3)
    cfg, twin = FenepmpJBP(mesh, lambda_d=0.1, We=0.5), FenepmpJBP(mesh, lambda_d=0.1, We=0.5)
    cfg.t = twin.t = 0.45
    ns = _jbp_runtime(cfg, mesh, dict(lambda_d="lambda_d", b="b_fene", A_0="a0"))
    src = open(ts.REF4).read().splitlines()
    for a, b in ((406, 407), (413, 413), (415, 415)):
        exec(textwrap.dedent("\n".join(src[a - 1:b])), ns)
    body = "if True:\n" + "\n".join(src[454:690]) + "\n"
    for step in range(K):
        exec(body, ns)
        ns["t"] = _number(ns["t"])
        r = twin.step()
        for key, got in (("E_k", ns["E_k"]), ("E_e", ns["E_e"]), ("torque", ns["y1"][-1]), ("F_x", ns["zx1"][-1]),
                         ("F_y", ns["z1"][-1])):
            scale = max(abs(r["F_x"]), abs(r["F_y"])) if key in ("F_x", "F_y") else abs(r[key])
            assert abs(got - r[key]) <= 1e-8 * max(scale, 1e-12), (step, key)
    _jbp_compare(ns, twin, K, ("w1", "p1", "tau1_vec", "T1", "rho1", "w0", "rho0", "T0", "ws", "rho12"), ())
    assert lc.rel_err(ns["kapp"].fn.vec, twin.kapp.vec) < 1e-7


@pytest.mark.parametrize("which", ["main", "prepass", "true_ewm"])
def test_config3_loop_body_text_runs_on_the_functors_for_k_steps(which):
    """incompressible_benchmarkEWM.py:466-782 (main pass: ramped spin, stress step behind its `if`; pre-pass: un-ramped spin)
    and comp_ewm_jpb.py:422-732 (the true extended White-Metzner loop) -- K steps from the text on the functors"""
    if not os.path.exists(ts.REF3):
        pytest.skip("reference checkout not present")
    import textwrap
    from oracle import configs_ewm as ce
    K = 3
    mesh = fem.annulus_mesh(12, 3)
    if which == "true_ewm":
        script, pre, loop = os.path.join(os.path.dirname(ts.REF3), "comp_ewm_jpb.py"), ((368, 368), (376, 377), (381, 382), (386, 389)), (422, 732)
        cfg, twin = ce.CompEwmJBP(mesh, dt=0.002), ce.CompEwmJBP(mesh, dt=0.002)
        prepass = False
    else:
        prepass = which == "prepass"
        script, pre, loop = ts.REF3, ((402, 402), (410, 411), (415, 416), (422, 422)), (466, 782)
        kw = dict(prepass=prepass, Ma=0.05, betav=0.6, A_0=0.4, k_ewm=0.7, B=0.2, dt=0.002)
        cfg, twin = ce.EwmJBP(mesh, **kw), ce.EwmJBP(mesh, **kw)
    cfg.t = twin.t = 0.45
    ns = _jbp_runtime(cfg, mesh, dict(A_0="a0", al="al1"))
    P = symbolic.Parameter
    ns.update(k_ewm=P("k_ewm", cfg.k_ewm), B=P("b_ewm", cfg.B), adaptive_loop=not prepass, err_count=0, max_refine=2)
    src = open(script).read().splitlines()
    for a, b in pre:
        exec(textwrap.dedent("\n".join(src[a - 1:b])), ns)
    raw = "\n".join(src[loop[0] - 1:loop[1]])
    indent = " " * (len(raw) - len(raw.lstrip(" ")))                  # the two scripts nest their loops differently
    body = "for once_ in (0,):\n" + raw + "\n" + indent + "completed_ = True\n"
    for step in range(K):
        ns["completed_"] = False
        exec(body, ns)
        assert ns["completed_"]
        ns["t"] = _number(ns["t"])
        r = twin.step()
        for key, got in (("E_k", ns["E_k"]), ("E_e", ns["E_e"]), ("torque", ns["y1"][-1]), ("F_y", ns["z1"][-1])):
            scale = max(abs(r["F_x"]), abs(r["F_y"])) if key == "F_y" else abs(r[key])
            floor = 1e-4 if key == "E_e" else 1e-12    # E_e = int(tr(tau) - 2) cancels to ~1e-10 at the start of the pre-pass: 1e-12 absolute
            assert abs(got - r[key]) <= 1e-8 * max(scale, floor), (step, key)
    _jbp_compare(ns, twin, K, ("w1", "p1", "tau1_vec", "T1", "rho1", "w0", "rho0", "T0", "w12", "ws"), ())


def test_incompressible_dgp_loop_body_text_runs_on_the_functors_for_k_steps():
    """buoyancy-driven-flow/incompressible_dgp.py:358-535 -- LPS block with absolute(kapp), `iter > 1` roll, two momentum
    systems, the singular pure-Neumann pressure system, velocity update, stress (behind `if betav < 0.99`), temperature,
    energies, Nusselt number: K steps from the text on the functors"""
    if not os.path.exists(ts.REF6):
        pytest.skip("reference checkout not present")
    import textwrap
    import fenics_b200.shims.buoyancy_driven_flow.fenics_fem as ff
    from oracle import configs_dgp as cd
    K = 3
    mesh = dgp_mesh(6)
    cfg, twin = cd.IncompDGP(mesh), cd.IncompDGP(mesh)
    cfg.t = twin.t = 1.4
    fields = dict(w0="W", w12="W", ws="W", w1="W", p0="Q", p1="Q", rho0="Q", rho1="Q", tau0_vec="Zc", tau1_vec="Zc", T0="Q",
                  T1="Q", thetar="Q")
    names = dict(Ra="Ra", Pr="Pr", We="We", dt="dt", betav="betav", th="th", Vh="Vh", Bi="Bi", c1="c1", c2="c2", T_0="T0",
                 T_h="th_hot")
    ns = _emu_runtime(ff, cfg, _plan(mesh, cfg), fields, names)
    ns["thetar"].slot = -1

    def solve(A, x, b, *a):
        singular = abs(A @ np.ones(A.shape[0])).max() < 1e-12 * abs(A).max()
        x.fn.vec[:] = cd.solve_weighted_gauge(A, b, x.fn.vec) if singular else fem.solve_lu(A, b)
    ns.update(solve=solve, solvertau=_EmuSolver(solve), al=0.0, f=symbolic.as_vector([0.0, -1.0]), k=symbolic.as_vector([0.0, 1.0]),
              bcT=cfg.bcT, ramp_function=_Setter(cfg._tt))
    ns["thetal"] = ns["T"] / (ns["T_h"] - ns["T_0"])
    ns["theta0"] = (ns["T0"] - ns["T_0"]) / (ns["T_h"] - ns["T_0"])
    src = open(ts.REF6).read().splitlines()
    exec(textwrap.dedent("\n".join(src[325:329])), ns)
    body = "if True:\n" + "\n".join(src[357:535]) + "\n"
    for _ in range(K):
        exec(body, ns)
        ns["t"] = _number(ns["t"])
        r = twin.step()
        for key in ("E_k", "E_e", "Nus"):
            assert abs(ns[key] - r[key]) <= 1e-8 * max(abs(r[key]), 1e-4 if key == "E_e" else 1e-12), key
    assert abs(ns["t"] - twin.t) < 1e-14 and ns["iter"] == K
    for name, ref in (("w1", twin.w1), ("p1", twin.p1), ("tau1_vec", twin.tau1_vec), ("T1", twin.T1), ("w0", twin.w0),
                      ("T0", twin.T0), ("w12", twin.w12), ("ws", twin.ws)):
        assert lc.rel_err(ns[name].fn.vec, ref.vec) < 1e-8, name
    assert lc.rel_err(ns["kapp"].fn.vec, twin.kapp.vec) < 1e-7
    assert np.abs(twin.w1.vec).max() > 1e-6
