"""Form recognition: makes the scripts' literal `assemble(a1)` work.

The reference builds every UFL form inside its time loop and hands it to dolfin (`A1 = assemble(a1)`,
lid-driven-cavity/incomp_LDC.py:317-331).  The GPU path has a hand-written functor per form (fx_form ids), so what is
needed between the two is not a form compiler but a form RECOGNISER: this module is a small symbolic tensor algebra
with UFL's operator names (grad, nabla_grad, div, dot, inner, transpose, as_matrix, Identity, `*dx`, `*ds(i)` ...) whose
expression trees have a canonical structural signature.  The registry (`fenics_b200/form_registry.py`) executes the
scripts' own form-building statements once on template objects and records signature -> fx_form id; a user form
with the same structure is then assembled by that functor, its Functions bound to the functor's field slots in order
of first appearance and its `Parameter`s written to the parameter table by name.  A form the library has no functor
for raises `UnknownForm` naming what it is missing -- never a silent fallback.

Scalars that parametrise the forms (Re, We, dt, betav ...) are `Parameter(name, value)` objects: the scripts hold them
as Python floats, which Python would fold into anonymous numbers before any recogniser sees them; wrapping them is
the one change a maintainer makes to a script's set-up section.  This module is pure Python and
imports neither torch nor the CUDA library.
"""
import numbers

import numpy as np


class UnknownForm(KeyError):
    pass


def _num(x):
    return repr(float(x))


class E:
    """expression node: op, children, tensor shape"""
    __slots__ = ("op", "a", "shape", "data")
    __array_priority__ = 1000  # numpy scalars defer to our reflected operators

    def __init__(self, op, a=(), shape=(), data=None):
        self.op, self.a, self.shape, self.data = op, tuple(a), tuple(shape), data

    # ---- signature ----------------------------------------------------------------------------------
    def sig(self, ctx):
        if self.op == "num":
            return _num(self.data)
        if self.op == "par":
            ctx.params[self.data.name] = self.data
            return "P:" + self.data.name
        if self.op == "arg":
            number, space, comps, space_obj = self.data
            ctx.args.add(number)
            ctx.spaces.add(space)
            if space_obj is not None:
                ctx.space_objs.append(space_obj)
            return "A%d:%s%s" % (number, space, list(comps))
        if self.op == "coef":
            fn, comps = self.data
            if id(fn) not in ctx.coef_index:
                ctx.coef_index[id(fn)] = len(ctx.coefs)
                ctx.coefs.append(fn)
            return "C%d:%s%s" % (ctx.coef_index[id(fn)], fn.function_space().name, list(comps))
        if self.op in ("h", "n"):
            return self.op
        inner_ = ",".join(c.sig(ctx) for c in self.a)
        extra = "" if self.data is None else "[%s]" % (self.data,)
        return "%s%s(%s)" % (self.op, extra, inner_)

    # ---- operators ------------------------------------------------------------------------------------
    def __add__(self, o):
        o = as_expr(o)
        if o.op == "num" and o.data == 0.0 and self.shape != ():
            return self
        _same(self, o)
        return E("add", (self, o), self.shape)

    def __radd__(self, o):
        return as_expr(o) + self

    def __sub__(self, o):
        o = as_expr(o)
        _same(self, o)
        return E("sub", (self, o), self.shape)

    def __rsub__(self, o):
        return as_expr(o) - self

    def __neg__(self):
        return E("neg", (self,), self.shape)

    def __abs__(self):  # Tr(tau) = abs(tau[0,0] + tau[1,1]) (fenics_base.py:246-247)
        return E("abs", (self,), self.shape)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        return _numpy_ufunc(ufunc, method, inputs)

    def __pos__(self):  # the scripts write `... + \ + inner(...)*dx` (comp_LDC_mesh_convergence.py:432-433)
        return self

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([(self, o.kind, o.subdomain)])
        if isinstance(o, Form):
            return o.__rmul__(self)
        o = as_expr(o)
        if self.shape == () or o.shape == ():
            return E("mul", (self, o), self.shape or o.shape)
        if len(self.shape) == 2 and len(o.shape) == 1:  # UFL: matrix * vector
            return dot(self, o)
        if len(self.shape) == 2 and len(o.shape) == 2:
            return dot(self, o)
        raise TypeError("unsupported product of shapes %s and %s" % (self.shape, o.shape))

    def __rmul__(self, o):
        if isinstance(o, Form):
            return NotImplemented
        return as_expr(o) * self

    def __truediv__(self, o):
        o = as_expr(o)
        if o.shape != ():
            raise TypeError("division by a non-scalar")
        return E("div", (self, o), self.shape)

    def __rtruediv__(self, o):
        return as_expr(o) / self

    def __pow__(self, o):
        return E("pow", (self, as_expr(o)), self.shape)

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        if len(idx) > len(self.shape) or any(not isinstance(i, (int, np.integer)) for i in idx):
            raise IndexError("fixed integer indices only")
        for i, n in zip(idx, self.shape):
            if not 0 <= i < n:
                raise IndexError(idx)
        return E("idx", (self,), self.shape[len(idx):], tuple(int(i) for i in idx))

    def __len__(self):
        if not self.shape:
            raise TypeError("scalar expression has no len()")
        return self.shape[0]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    @property
    def T(self):
        return transpose(self)

    # ---- constant trees are numbers -------------------------------------------------------------------
    # With dt a Parameter the scripts' own bookkeeping -- `t += dt` ... `if t > 0.5:` (incomp_LDC.py:429, :437),
    # `(Tf-t)/dt * elapsed` (:284) -- builds trees of numbers and Parameters only: they evaluate where a number is needed.
    def __float__(self):
        return _const_value(self)

    def __lt__(self, o):
        return float(self) < float(o)

    def __gt__(self, o):
        return float(self) > float(o)

    def __le__(self, o):
        return float(self) <= float(o)

    def __ge__(self, o):
        return float(self) >= float(o)


def _const_value(e):
    if e.op == "num":
        return float(e.data)
    if e.op == "par":
        return float(e.data.value)
    if e.op in ("add", "sub", "mul", "div", "neg", "pow", "abs", "sqrt") and e.shape == ():
        a = [_const_value(c) for c in e.a]
        if e.op == "add":
            return a[0] + a[1]
        if e.op == "sub":
            return a[0] - a[1]
        if e.op == "mul":
            return a[0] * a[1]
        if e.op == "div":
            return a[0] / a[1]
        if e.op == "neg":
            return -a[0]
        if e.op == "pow":
            return a[0] ** a[1]
        if e.op == "abs":
            return abs(a[0])
        return a[0] ** 0.5
    raise TypeError("expression with fields or arguments (%s) used as a number" % e.op)


def _same(a, b):
    if a.shape != b.shape:
        raise TypeError("shape mismatch %s vs %s" % (a.shape, b.shape))


class Parameter(E):
    """named scalar of a run (Re, We, dt, betav ...): a symbolic leaf that remembers its parameter-table slot by
    name and still behaves as a number where the scripts use it as one (float(), comparisons, printing)."""
    __slots__ = ("name", "value")

    def __init__(self, name, value):
        E.__init__(self, "par", (), (), self)
        self.name, self.value = name, float(value)

    def __float__(self):
        return self.value

    def __repr__(self):
        return repr(self.value)

    def __lt__(self, o):
        return self.value < float(o)

    def __gt__(self, o):
        return self.value > float(o)

    def __le__(self, o):
        return self.value <= float(o)

    def __ge__(self, o):
        return self.value >= float(o)


def _numpy_ufunc(ufunc, method, inputs):
    """the scripts apply numpy to UFL objects -- `res_orth_norm = np.power(res_orth_norm_sq, 0.5)` (incomp_LDC.py:307),
    `np.power(u[0]*u[0]+u[1]*u[1], 0.5)` (magnitude): the same operators, symbolically"""
    name = getattr(ufunc, "__name__", "")
    if method == "__call__" and name == "power" and len(inputs) == 2:
        return as_expr(inputs[0]) ** inputs[1]
    if method == "__call__" and name in ("absolute", "fabs") and len(inputs) == 1:
        return abs(as_expr(inputs[0]))
    if method == "__call__" and name == "sqrt" and len(inputs) == 1:
        return as_expr(inputs[0]) ** 0.5
    return NotImplemented


class Operand:
    """mixin for Function-like objects: inside a form they behave as their coefficient expression"""

    def _sym_expr(self):
        return coefficient(self)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        return _numpy_ufunc(ufunc, method, inputs)

    def __abs__(self):
        return abs(as_expr(self))

    def __add__(self, o):
        return as_expr(self) + o

    def __radd__(self, o):
        return o + as_expr(self)

    def __sub__(self, o):
        return as_expr(self) - o

    def __rsub__(self, o):
        return o - as_expr(self)

    def __neg__(self):
        return -as_expr(self)

    def __mul__(self, o):
        return as_expr(self) * o

    def __rmul__(self, o):
        return o * as_expr(self)

    def __truediv__(self, o):
        return as_expr(self) / o

    def __pow__(self, o):
        return as_expr(self) ** o

    def __getitem__(self, i):
        return as_expr(self)[i]


def as_expr(x):
    if isinstance(x, E):
        return x
    if isinstance(x, (numbers.Real, np.floating, np.integer)):
        return E("num", (), (), float(x))
    if hasattr(x, "_sym_expr"):
        return x._sym_expr()
    raise TypeError("cannot use %r in a form" % (type(x).__name__,))


# ---- leaves -------------------------------------------------------------------------------------------
def argument(number, space, comps, space_obj=None):
    """test (number 0) / trial (number 1) function of the space NAMED `space`, scalar components `comps`; space_obj: the
    FunctionSpace (leads assemble() to the device plan when a form has no coefficient)"""
    comps = tuple(comps)
    return E("arg", (), (len(comps),) if len(comps) > 1 else (), (number, space, comps, space_obj))


def coefficient(fn, comps=None):
    """a (sub-)Function: comps = scalar components of its space it covers (None: all)"""
    sp = fn.function_space()
    if comps is None:
        comps = tuple(range(len(sp.comps)))
    comps = tuple(comps)
    return E("coef", (), (len(comps),) if len(comps) > 1 else (), (fn, comps))


def Constant(value):
    """dolfin Constant inside a form: a literal number / vector of numbers (`Constant((1.0, 0.0))`, fenepmp_jbp.py:647)"""
    if isinstance(value, (tuple, list)):
        return as_vector([float(v) for v in value])
    return as_expr(float(value))


def CellDiameter(mesh=None):
    return E("h")


def FacetNormal(mesh=None):
    return E("n", (), (2,))


def Identity(n):
    return E("identity", (), (n, n), int(n))


# ---- operators with UFL's names and index conventions ---------------------------------------------------
def grad(e):
    e = as_expr(e)
    return E("grad", (e,), e.shape + (2,))


def nabla_grad(e):
    e = as_expr(e)
    return E("nabla_grad", (e,), (2,) + e.shape)


def div(e):
    e = as_expr(e)
    if not e.shape or e.shape[-1] != 2:
        raise TypeError("div of shape %s" % (e.shape,))
    return E("divop", (e,), e.shape[:-1])


def dot(a, b):
    a, b = as_expr(a), as_expr(b)
    if not a.shape or not b.shape or a.shape[-1] != b.shape[0]:
        raise TypeError("dot of shapes %s and %s" % (a.shape, b.shape))
    return E("dot", (a, b), a.shape[:-1] + b.shape[1:])


def inner(a, b):
    a, b = as_expr(a), as_expr(b)
    _same(a, b)
    return E("inner", (a, b), ())


def outer(a, b):
    a, b = as_expr(a), as_expr(b)
    return E("outer", (a, b), a.shape + b.shape)


def transpose(e):
    e = as_expr(e)
    if len(e.shape) != 2:
        raise TypeError("transpose of shape %s" % (e.shape,))
    return E("T", (e,), e.shape[::-1])


def tr(e):
    e = as_expr(e)
    return E("tr", (e,), ())


def sym(e):
    e = as_expr(e)
    return E("sym", (e,), e.shape)


def sqrt(e):
    return E("sqrt", (as_expr(e),), ())


def as_vector(items):
    items = [as_expr(i) for i in items]
    return E("vec", items, (len(items),))


def as_matrix(rows):
    rows = [[as_expr(i) for i in r] for r in rows]
    return E("mat", [i for r in rows for i in r], (len(rows), len(rows[0])), len(rows[0]))


# ---- measures and forms -----------------------------------------------------------------------------------
class Measure:
    def __init__(self, kind, subdomain=None):
        self.kind, self.subdomain = kind, subdomain

    def __call__(self, subdomain):
        return Measure(self.kind, int(subdomain))

    def __rmul__(self, e):
        return as_expr(e) * self


dx = Measure("dx")
ds = Measure("ds")


class Form:
    """sum of integrals; supports the arithmetic the scripts do on forms (a += th*DEVSSl, L -= ...; lhs / rhs below)"""

    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, o):
        if isinstance(o, numbers.Real) and o == 0:
            return self
        if not isinstance(o, Form):
            return NotImplemented
        return Form(self.integrals + o.integrals)

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-1.0) * o

    def __neg__(self):
        return (-1.0) * self

    def __pos__(self):
        return self

    def __rmul__(self, c):
        c = as_expr(c)
        if c.shape != ():
            raise TypeError("only scalars scale a form")
        return Form([(E("mul", (c, e), e.shape), k, s) for e, k, s in self.integrals])

    def __mul__(self, c):
        return self.__rmul__(c)

    def signature(self):
        ctx = _Ctx()
        parts = ["%s%s:%s" % (k, "" if s is None else "(%d)" % s, e.sig(ctx)) for e, k, s in self.integrals]
        return "|".join(parts), ctx


class _Ctx:
    def __init__(self):
        self.params, self.args, self.spaces, self.coefs, self.coef_index = {}, set(), set(), [], {}
        self.space_objs = []


# ---- lhs / rhs: the scripts that write one residual F(u; v) = 0 and let UFL split it --------------------------
# (comp_LDC_mesh_convergence.py:439-440 `a1 = lhs(Fu12); L1 = rhs(Fu12)`).  The split is structural: every node is
# separated into the part that contains the trial function (linearly) and the part that does not.
_LINEAR_UNARY = ("grad", "nabla_grad", "divop", "T", "tr", "sym", "idx", "neg")
_BILINEAR = ("mul", "dot", "inner", "outer")


def _zero_like(e):
    z = E("num", data=0.0)
    return z if e.shape == () else E("zero", (), e.shape)


def _split_trial(e):
    """(part with the trial function, part without); either may be None (= zero)"""
    if e.op == "arg":
        return (e, None) if e.data[0] == 1 else (None, e)
    if not e.a:
        return None, e
    if e.op in ("add", "sub"):
        (aw, ao), (bw, bo) = _split_trial(e.a[0]), _split_trial(e.a[1])

        def comb(x, y):
            if x is None and y is None:
                return None
            if y is None:
                return x
            if x is None:
                return y if e.op == "add" else E("neg", (y,), y.shape)
            return E(e.op, (x, y), e.shape)
        return comb(aw, bw), comb(ao, bo)
    if e.op in _LINEAR_UNARY:
        w, o = _split_trial(e.a[0])
        mk = lambda x: None if x is None else E(e.op, (x,), e.shape, e.data)  # noqa: E731
        return mk(w), mk(o)
    if e.op in ("vec", "mat"):
        parts = [_split_trial(c) for c in e.a]
        if all(w is None for w, _ in parts):
            return None, e
        mk = lambda k: E(e.op, [p[k] if p[k] is not None else _zero_like(c) for p, c in zip(parts, e.a)], e.shape, e.data)  # noqa: E731
        return mk(0), (None if all(o is None for _, o in parts) else mk(1))
    if e.op in _BILINEAR:
        (aw, ao), (bw, bo) = _split_trial(e.a[0]), _split_trial(e.a[1])
        if aw is not None and bw is not None:
            raise ValueError("lhs/rhs: the form is not linear in the trial function")
        mk = lambda x, y: E(e.op, (x, y), e.shape, e.data)  # noqa: E731
        terms = [mk(x, y) for x, y in ((aw, bo), (ao, bw)) if x is not None and y is not None]
        w = None if not terms else terms[0] if len(terms) == 1 else E("add", terms, e.shape)
        return w, (mk(ao, bo) if ao is not None and bo is not None else None)
    if e.op == "div":
        (aw, ao), (bw, _) = _split_trial(e.a[0]), _split_trial(e.a[1])
        if bw is not None:
            raise ValueError("lhs/rhs: trial function in a denominator")
        mk = lambda x: None if x is None else E("div", (x, e.a[1]), e.shape)  # noqa: E731
        return mk(aw), mk(ao)
    if any(_split_trial(c)[0] is not None for c in e.a):  # pow, sqrt ...
        raise ValueError("lhs/rhs: trial function inside the non-linear operator %r" % e.op)
    return None, e


def lhs(form):
    """the bilinear part of F (UFL's lhs)"""
    out = []
    for e, k, s in form.integrals:
        w, _ = _split_trial(e)
        if w is not None:
            out.append((w, k, s))
    return Form(out)


def rhs(form):
    """minus the part of F without the trial function (UFL's rhs: F = a - L)"""
    out = []
    for e, k, s in form.integrals:
        _, o = _split_trial(e)
        if o is not None:
            out.append((E("neg", (o,), o.shape), k, s))
    return Form(out)


def system(form):
    return lhs(form), rhs(form)


# ---- registry ----------------------------------------------------------------------------------------------
_REGISTRY = {}


def register(form, form_id, slots, what, derived=None):
    """form: a Form built from template objects; slots: fx_field slot of every coefficient in order of first
    appearance; what: the script statement it restates (file:line) for error messages; derived: {parameter name:
    f(values by name)} for a functor whose parameter slot holds a COMBINATION of the script's scalars (comp_LDC.py writes
    `Re*conv*dot(u, nabla_grad(u))` where the functor multiplies by its `conv` slot)"""
    sig, ctx = form.signature()
    if len(ctx.coefs) != len(slots):
        raise ValueError("%s: %d coefficients in the form, %d slots given" % (what, len(ctx.coefs), len(slots)))
    # two scripts may write the SAME form (the stress matrix of configs 1 and 2 is one operator, served by one functor
    # under two ids): the first registration stands
    fid = form_id if isinstance(form_id, tuple) else int(form_id)   # ("expr", id): a DG0 projection kernel (fx_project_dg0)
    _REGISTRY.setdefault(sig, (fid, tuple(int(s) for s in slots), what, dict(derived or {})))


def recognise(form):
    """-> (fx_form id, [(slot, Function)], {parameter name: Parameter}, description, [FunctionSpace of the arguments]);
    raises UnknownForm"""
    from . import form_registry  # noqa: F401  (fills _REGISTRY on first use)
    form_registry.load()
    sig, ctx = form.signature()
    hit = _REGISTRY.get(sig)
    if hit is None:
        raise UnknownForm("no hand-written functor matches this form (rank %d on %s, %d integrals, %d coefficients); "
                          "the library recognises %d forms of the reference scripts -- see fenics_b200/form_registry.py"
                          % (len(ctx.args), sorted(ctx.spaces), len(form.integrals), len(ctx.coefs), len(_REGISTRY)))
    fid, slots, what, derived = hit
    params = dict(ctx.params)
    if derived:
        values = {k: float(v.value) for k, v in params.items()}
        for name, fn in derived.items():
            params[name] = Parameter(name, fn(values))
    return fid, list(zip(slots, ctx.coefs)), params, what, ctx.space_objs
