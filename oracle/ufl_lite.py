"""ORACLE (test infrastructure, NOT product code) -- a tiny UFL-semantics expression layer.

parity unpinned: the reference (ATMackay/fenics) ships no golden vectors for this path and
dolfin/UFL/FFC 2019.1.0 are not installable here; this file restates the published UFL
semantics the reference's call sites rely on, and is pinned only by construction-independent
checks (sympy-exact element tensors, identities; see tests/test_oracle_*.py).

What it restates (third-party, un-vendored: UFL 2019.1.0 -- pinned by fenics-install.sh:28-40):
  * tensor algebra used by the reference helpers (lid-driven-cavity/fenics_fem.py:289-314):
    grad (derivative index LAST), nabla_grad (derivative index FIRST), div (contracts the last
    index), dot (last index of a with first index of b), inner (full contraction), `*`
    (scalar scaling; matrix*vector and matrix*matrix are dot), .T, indexing, as_vector/as_matrix
  * literal-zero folding: python 0/0.0 times anything is Zero (UFL FloatValue.__new__)
  * lhs()/rhs() splitting of a residual form into its trial-linear and trial-free parts
  * total polynomial degree estimation (UFL SumDegreeEstimator) used to pick quadrature

Evaluation is numerical: every node evaluates to an array of shape (ne, nq) + tensor_shape over
(entities, quadrature points); arguments (test/trial functions) are evaluated as unit "slot"
indicators (component c: value / d/dx / d/dy) so that a bilinear integrand yields the pointwise
coefficient tensor C[s, t] directly (see oracle/fem.py:assemble).
"""
import numpy as np

_LET = "abcdfghijklmnoprstuvwxyz"


def as_expr(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, (int, float, np.integer, np.floating)):
        return Zero(()) if float(x) == 0.0 else Const(float(x))
    if isinstance(x, (list, tuple, np.ndarray)):
        arr = np.asarray(x, dtype=float)
        return Const(arr)
    raise TypeError("cannot convert %r to Expr" % (x,))


class Expr:
    __array_priority__ = 1000
    __array_ufunc__ = None  # make numpy scalars defer to our reflected operators
    shape = ()
    deps = frozenset()

    # ---- operators -------------------------------------------------------------------------
    def __add__(self, o):
        if isinstance(o, Form):
            return NotImplemented
        return Sum.make(self, as_expr(o))

    def __radd__(self, o):
        return Sum.make(as_expr(o), self)

    def __sub__(self, o):
        return Sum.make(self, Neg.make(as_expr(o)))

    def __rsub__(self, o):
        return Sum.make(as_expr(o), Neg.make(self))

    def __neg__(self):
        return Neg.make(self)

    def __pos__(self):
        return self

    def __abs__(self):
        return Math("abs", self)

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([Integral(self, o.kind, o.subdomain_id)])
        return _mult(self, as_expr(o))

    def __rmul__(self, o):
        return _mult(as_expr(o), self)

    def __truediv__(self, o):
        return Division.make(self, as_expr(o))

    def __rtruediv__(self, o):
        return Division.make(as_expr(o), self)

    def __pow__(self, p):
        return Power.make(self, p)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Indexed.make(self, tuple(int(i) for i in idx))

    def __len__(self):
        if len(self.shape) != 1:
            raise TypeError("len() of non-vector expression")
        return self.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    @property
    def T(self):
        return Transposed.make(self)

    def dx(self, i):
        return Dx.make(self, int(i))

    # ---- interface -------------------------------------------------------------------------
    def ev(self, ctx):
        if not self.deps:
            key = (id(self), 0)
            r = ctx.memo.get(key)
            if r is None:
                r = self._ev(ctx)
                ctx.memo[key] = r
                ctx.keep.append(self)
            return r
        return self._ev(ctx)

    def gr(self, ctx):
        if not self.deps:
            key = (id(self), 1)
            r = ctx.memo.get(key)
            if r is None:
                r = self._gr(ctx)
                ctx.memo[key] = r
                ctx.keep.append(self)
            return r
        return self._gr(ctx)

    def _gr(self, ctx):
        raise NotImplementedError("gradient of %s" % type(self).__name__)

    def deg(self):
        raise NotImplementedError(type(self).__name__)

    def split(self, num):
        """-> (part free of argument `num`, part linear in argument `num`); None means zero."""
        raise NotImplementedError(type(self).__name__)


def _is_zero(e):
    return isinstance(e, Zero)


def _is_scalar_const(e):
    return isinstance(e, Const) and e.shape == () and e._deg == 0


def _zeros(ctx, shape):
    return np.zeros((1, 1) + tuple(shape))


class Zero(Expr):
    def __init__(self, shape=()):
        self.shape = tuple(shape)

    def _ev(self, ctx):
        return _zeros(ctx, self.shape)

    def _gr(self, ctx):
        return _zeros(ctx, self.shape + (2,))

    def deg(self):
        return 0

    def split(self, num):
        return None, None


class Const(Expr):
    """Python literal / dolfin Constant / constant Expression (degree 0 in UFL's estimate)."""

    def __init__(self, value, degree=0):
        self.value = np.asarray(value, dtype=float)
        self.shape = self.value.shape
        self._deg = degree

    def _ev(self, ctx):
        return self.value.reshape((1, 1) + self.shape)

    def _gr(self, ctx):
        return _zeros(ctx, self.shape + (2,))

    def deg(self):
        return self._deg

    def split(self, num):
        return self, None


def Constant(v):
    return as_expr(v) if np.ndim(v) == 0 else Const(np.asarray(v, dtype=float))


def Identity(n):
    return Const(np.eye(n))


class Coef(Expr):
    """A Function (or a .split()/.sub() VIEW of one: `comps` selects components of the parent
    space; the parent vector is read at evaluation time, so views alias like dolfin's)."""

    def __init__(self, fn, comps, scalar):
        self.fn, self.comps, self.scalar = fn, tuple(comps), scalar
        self.shape = () if scalar else (len(comps),)

    def _ev(self, ctx):
        v = ctx.coef_value(self.fn, self.comps)  # (ne,nq,ncomp)
        return v[..., 0] if self.scalar else v

    def _gr(self, ctx):
        g = ctx.coef_grad(self.fn, self.comps)  # (ne,nq,ncomp,2)
        return g[..., 0, :] if self.scalar else g

    def deg(self):
        # sub-functions carry their own sub-element degree (dolfin Function.sub(i))
        return max(self.fn.space.comp_degree(c) for c in self.comps)

    def split(self, num):
        return self, None


class Arg(Expr):
    """Test (number 0) / trial (number 1) function, or components of a mixed one."""

    def __init__(self, space, number, comps, scalar):
        self.space, self.number, self.comps, self.scalar = space, number, tuple(comps), scalar
        self.shape = () if scalar else (len(comps),)
        self.deps = frozenset([number])

    def _ev(self, ctx):
        val, _ = ctx.args[self.number]
        v = val[list(self.comps)].reshape((1, 1, len(self.comps)))
        return v[..., 0] if self.scalar else v

    def _gr(self, ctx):
        _, grad = ctx.args[self.number]
        g = grad[list(self.comps)].reshape((1, 1, len(self.comps), 2))
        return g[..., 0, :] if self.scalar else g

    def deg(self):
        # components of ONE Argument on the mixed element: degree = max sub-degree
        return self.space.degree()

    def split(self, num):
        return (None, self) if num == self.number else (self, None)


class CellDiameter(Expr):
    def __init__(self, mesh=None):
        pass

    def _ev(self, ctx):
        return ctx.h[:, None]

    def _gr(self, ctx):
        return _zeros(ctx, (2,))

    def deg(self):
        return 0

    def split(self, num):
        return self, None


class FacetNormal(Expr):
    shape = (2,)

    def __init__(self, mesh=None):
        pass

    def _ev(self, ctx):
        if ctx.normal is None:
            raise ValueError("FacetNormal in a cell integral")
        return ctx.normal[:, None, :]

    def _gr(self, ctx):
        return _zeros(ctx, (2, 2))

    def deg(self):
        return 0

    def split(self, num):
        return self, None


def _padd(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return Sum.make(a, b)


class Sum(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        self.shape = a.shape
        self.deps = a.deps | b.deps

    @staticmethod
    def make(a, b):
        if a.shape != b.shape:
            raise ValueError("shape mismatch in sum: %s vs %s" % (a.shape, b.shape))
        if _is_zero(a):
            return b
        if _is_zero(b):
            return a
        if _is_scalar_const(a) and _is_scalar_const(b):  # UFL folds ScalarValue + ScalarValue
            return as_expr(float(a.value) + float(b.value))
        return Sum(a, b)

    def _ev(self, ctx):
        return self.a.ev(ctx) + self.b.ev(ctx)

    def _gr(self, ctx):
        return self.a.gr(ctx) + self.b.gr(ctx)

    def deg(self):
        return max(self.a.deg(), self.b.deg())

    def split(self, num):
        a0, a1 = self.a.split(num)
        b0, b1 = self.b.split(num)
        return _padd(a0, b0), _padd(a1, b1)


class Neg(Expr):
    def __init__(self, a):
        self.a = a
        self.shape = a.shape
        self.deps = a.deps

    @staticmethod
    def make(a):
        if _is_zero(a):
            return a
        if isinstance(a, Neg):
            return a.a
        if _is_scalar_const(a):
            return as_expr(-float(a.value))
        return Neg(a)

    def _ev(self, ctx):
        return -self.a.ev(ctx)

    def _gr(self, ctx):
        return -self.a.gr(ctx)

    def deg(self):
        return self.a.deg()

    def split(self, num):
        a0, a1 = self.a.split(num)
        return (None if a0 is None else Neg.make(a0)), (None if a1 is None else Neg.make(a1))


def _mult(a, b):
    """UFL `*`: scalar scaling, or dot for matrix*vector / matrix*matrix."""
    if a.shape == () or b.shape == ():
        return Product.make(a, b)
    if len(a.shape) == 2 and len(b.shape) in (1, 2):
        return Dot.make(a, b)
    raise ValueError("invalid ranks in product: %s * %s" % (a.shape, b.shape))


def _bilinear_split(make, a, b, num):
    a0, a1 = a.split(num)
    b0, b1 = b.split(num)
    if a1 is not None and b1 is not None:
        raise ValueError("expression is nonlinear in argument %d" % num)
    p0 = make(a0, b0) if (a0 is not None and b0 is not None) else None
    p1 = None
    if a1 is not None and b0 is not None:
        p1 = make(a1, b0)
    if a0 is not None and b1 is not None:
        p1 = _padd(p1, make(a0, b1))
    return p0, p1


class Product(Expr):
    """scalar * tensor (a is the scalar)."""

    def __init__(self, a, b):
        self.a, self.b = a, b
        self.shape = b.shape
        self.deps = a.deps | b.deps

    @staticmethod
    def make(a, b):
        if a.shape != ():
            a, b = b, a
        if a.shape != ():
            raise ValueError("Product needs a scalar operand")
        if _is_zero(a) or _is_zero(b):
            return Zero(b.shape)
        if _is_scalar_const(a) and _is_scalar_const(b):
            return as_expr(float(a.value) * float(b.value))
        return Product(a, b)

    def _ev(self, ctx):
        s = self.a.ev(ctx)
        return s.reshape(s.shape + (1,) * len(self.shape)) * self.b.ev(ctx)

    def _gr(self, ctx):
        s, sg = self.a.ev(ctx), self.a.gr(ctx)  # (e,q), (e,q,2)
        t, tg = self.b.ev(ctx), self.b.gr(ctx)  # (e,q,*S), (e,q,*S,2)
        n = len(self.shape)
        s_ = s.reshape(s.shape + (1,) * (n + 1))
        sg_ = sg.reshape(sg.shape[:2] + (1,) * n + (2,))
        return s_ * tg + sg_ * t[..., None]

    def deg(self):
        return self.a.deg() + self.b.deg()

    def split(self, num):
        return _bilinear_split(Product.make, self.a, self.b, num)


class Division(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        self.shape = a.shape
        self.deps = a.deps | b.deps

    @staticmethod
    def make(a, b):
        if b.shape != ():
            raise ValueError("division by non-scalar")
        if _is_zero(a):
            return a
        if _is_scalar_const(a) and _is_scalar_const(b):
            return as_expr(float(a.value) / float(b.value))
        return Division(a, b)

    def _ev(self, ctx):
        d = self.b.ev(ctx)
        return self.a.ev(ctx) / d.reshape(d.shape + (1,) * len(self.shape))

    def _gr(self, ctx):
        n = len(self.shape)
        d, dg = self.b.ev(ctx), self.b.gr(ctx)
        t, tg = self.a.ev(ctx), self.a.gr(ctx)
        d_ = d.reshape(d.shape + (1,) * (n + 1))
        dg_ = dg.reshape(dg.shape[:2] + (1,) * n + (2,))
        return tg / d_ - t[..., None] * dg_ / (d_ * d_)

    def deg(self):
        return self.a.deg() + self.b.deg()

    def split(self, num):
        b0, b1 = self.b.split(num)
        if b1 is not None:
            raise ValueError("argument in denominator")
        a0, a1 = self.a.split(num)
        return (None if a0 is None else Division.make(a0, self.b),
                None if a1 is None else Division.make(a1, self.b))


class Power(Expr):
    def __init__(self, a, p):
        self.a, self.p = a, p
        self.deps = a.deps

    @staticmethod
    def make(a, p):
        if a.shape != ():
            raise ValueError("power of non-scalar")
        if isinstance(p, Expr):
            raise NotImplementedError("expression exponent")
        p = float(p)
        if p == 0.0:
            return Const(1.0)  # UFL: f**0 -> IntValue(1)
        if p == 1.0:
            return a
        if _is_zero(a):
            return a
        if _is_scalar_const(a):
            return as_expr(float(a.value) ** p)
        return Power(a, p)

    def _ev(self, ctx):
        return self.a.ev(ctx) ** self.p

    def _gr(self, ctx):
        v = self.a.ev(ctx)
        with np.errstate(divide="ignore", invalid="ignore"):
            d = self.p * v ** (self.p - 1.0)
        return d[..., None] * self.a.gr(ctx)

    def deg(self):
        p = self.p
        if p >= 0 and float(p).is_integer():
            return self.a.deg() * int(p)
        return self.a.deg() + 2

    def split(self, num):
        if num in self.deps:
            raise ValueError("argument under a power")
        return self, None


class Math(Expr):
    _F = {"sqrt": (np.sqrt, lambda v: 0.5 / np.sqrt(v)),
          "tanh": (np.tanh, lambda v: 1.0 - np.tanh(v) ** 2),
          "exp": (np.exp, np.exp),
          "abs": (np.abs, np.sign)}

    def __init__(self, name, a):
        if a.shape != ():
            raise ValueError("math function of non-scalar")
        self.name, self.a = name, a
        self.deps = a.deps

    def _ev(self, ctx):
        return Math._F[self.name][0](self.a.ev(ctx))

    def _gr(self, ctx):
        with np.errstate(divide="ignore", invalid="ignore"):
            d = Math._F[self.name][1](self.a.ev(ctx))
        return d[..., None] * self.a.gr(ctx)

    def deg(self):
        return self.a.deg() if self.name == "abs" else self.a.deg() + 2

    def split(self, num):
        if num in self.deps:
            raise ValueError("argument under a math function")
        return self, None


def sqrt(a):
    return Math("sqrt", as_expr(a))


def tanh(a):
    return Math("tanh", as_expr(a))


def exp(a):
    return Math("exp", as_expr(a))


def ufl_abs(a):
    return Math("abs", as_expr(a))


class Dot(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        self.shape = a.shape[:-1] + b.shape[1:]
        self.deps = a.deps | b.deps
        ra, rb = len(a.shape), len(b.shape)
        la = _LET[:ra - 1] + "k"
        lb = "k" + _LET[ra - 1:ra - 1 + rb - 1]
        out = la[:-1] + lb[1:]
        self._s = "eq%s,eq%s->eq%s" % (la, lb, out)
        self._sga = "eq%sZ,eq%s->eq%sZ" % (la, lb, out)
        self._sgb = "eq%s,eq%sZ->eq%sZ" % (la, lb, out)

    @staticmethod
    def make(a, b):
        if not a.shape or not b.shape:
            return Product.make(a, b)
        if a.shape[-1] != b.shape[0]:
            raise ValueError("dot: dimension mismatch %s . %s" % (a.shape, b.shape))
        if _is_zero(a) or _is_zero(b):
            return Zero(a.shape[:-1] + b.shape[1:])
        return Dot(a, b)

    def _ev(self, ctx):
        return np.einsum(self._s, self.a.ev(ctx), self.b.ev(ctx))

    def _gr(self, ctx):
        return (np.einsum(self._sga, self.a.gr(ctx), self.b.ev(ctx))
                + np.einsum(self._sgb, self.a.ev(ctx), self.b.gr(ctx)))

    def deg(self):
        return self.a.deg() + self.b.deg()

    def split(self, num):
        return _bilinear_split(Dot.make, self.a, self.b, num)


class Inner(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        self.deps = a.deps | b.deps
        l = _LET[:len(a.shape)]
        self._s = "eq%s,eq%s->eq" % (l, l)
        self._sga = "eq%sZ,eq%s->eqZ" % (l, l)
        self._sgb = "eq%s,eq%sZ->eqZ" % (l, l)

    @staticmethod
    def make(a, b):
        if a.shape != b.shape:
            raise ValueError("inner: shape mismatch %s : %s" % (a.shape, b.shape))
        if _is_zero(a) or _is_zero(b):
            return Zero(())
        if a.shape == ():
            return Product.make(a, b)
        return Inner(a, b)

    def _ev(self, ctx):
        return np.einsum(self._s, self.a.ev(ctx), self.b.ev(ctx))

    def _gr(self, ctx):
        return (np.einsum(self._sga, self.a.gr(ctx), self.b.ev(ctx))
                + np.einsum(self._sgb, self.a.ev(ctx), self.b.gr(ctx)))

    def deg(self):
        return self.a.deg() + self.b.deg()

    def split(self, num):
        return _bilinear_split(Inner.make, self.a, self.b, num)


def dot(a, b):
    return Dot.make(as_expr(a), as_expr(b))


def inner(a, b):
    return Inner.make(as_expr(a), as_expr(b))


def _linear_split(make, a, num):
    a0, a1 = a.split(num)
    return (None if a0 is None else make(a0)), (None if a1 is None else make(a1))


class Indexed(Expr):
    def __init__(self, a, idx):
        self.a, self.idx = a, idx
        self.shape = a.shape[len(idx):]
        self.deps = a.deps

    @staticmethod
    def make(a, idx):
        if len(idx) > len(a.shape):
            raise IndexError("too many indices")
        for i, n in zip(idx, a.shape):
            if not 0 <= i < n:
                raise IndexError("index out of range")
        if _is_zero(a):
            return Zero(a.shape[len(idx):])
        if isinstance(a, ListTensor):
            r = a.items[idx[0]]
            return r if len(idx) == 1 else Indexed.make(r, idx[1:])
        return Indexed(a, idx)

    def _ev(self, ctx):
        return self.a.ev(ctx)[(slice(None), slice(None)) + self.idx]

    def _gr(self, ctx):
        return self.a.gr(ctx)[(slice(None), slice(None)) + self.idx]

    def deg(self):
        return self.a.deg()

    def split(self, num):
        return _linear_split(lambda x: Indexed.make(x, self.idx), self.a, num)


class ListTensor(Expr):
    def __init__(self, items):
        self.items = items
        self.shape = (len(items),) + items[0].shape
        d = frozenset()
        for it in items:
            d |= it.deps
        self.deps = d

    @staticmethod
    def make(items):
        items = [ListTensor.make(i) if isinstance(i, (list, tuple)) else as_expr(i) for i in items]
        sh = items[0].shape
        if any(i.shape != sh for i in items):
            raise ValueError("ragged list tensor")
        if all(_is_zero(i) for i in items):
            return Zero((len(items),) + sh)
        return ListTensor(items)

    def _bc(self, arrs):
        ne = max(a.shape[0] for a in arrs)
        nq = max(a.shape[1] for a in arrs)
        arrs = [np.broadcast_to(a, (ne, nq) + a.shape[2:]) for a in arrs]
        return np.stack(arrs, axis=2)

    def _ev(self, ctx):
        return self._bc([i.ev(ctx) for i in self.items])

    def _gr(self, ctx):
        return self._bc([i.gr(ctx) for i in self.items])

    def deg(self):
        return max(i.deg() for i in self.items)

    def split(self, num):
        parts = [i.split(num) for i in self.items]
        out = []
        for k in (0, 1):
            if all(p[k] is None for p in parts):
                out.append(None)
            else:
                out.append(ListTensor.make([Zero(self.items[0].shape) if p[k] is None else p[k]
                                            for p in parts]))
        return tuple(out)


def as_vector(items):
    return ListTensor.make(list(items))


def as_matrix(rows):
    return ListTensor.make([list(r) for r in rows])


class Transposed(Expr):
    def __init__(self, a):
        self.a = a
        self.shape = (a.shape[1], a.shape[0])
        self.deps = a.deps

    @staticmethod
    def make(a):
        if len(a.shape) != 2:
            raise ValueError("transpose of non-matrix")
        if _is_zero(a):
            return Zero((a.shape[1], a.shape[0]))
        return Transposed(a)

    def _ev(self, ctx):
        return np.swapaxes(self.a.ev(ctx), 2, 3)

    def _gr(self, ctx):
        return np.swapaxes(self.a.gr(ctx), 2, 3)

    def deg(self):
        return self.a.deg()

    def split(self, num):
        return _linear_split(Transposed.make, self.a, num)


def transpose(a):
    return Transposed.make(as_expr(a))


def _dred(d):
    # UFL SumDegreeEstimator._reduce_degree on affine simplices
    return max(d - 1, 0)


class Grad(Expr):
    """grad(f)[..., k] = d f[...] / d x_k  (derivative index last)."""

    def __init__(self, a):
        self.a = a
        self.shape = a.shape + (2,)
        self.deps = a.deps

    @staticmethod
    def make(a):
        if _is_zero(a) or isinstance(a, Const):
            return Zero(a.shape + (2,))
        return Grad(a)

    def _ev(self, ctx):
        return self.a.gr(ctx)

    def _gr(self, ctx):
        # second derivatives: only of a Function (view) itself -- what the reference's forms need
        # (grad of Dincomp(u0) inside div(phi_def(u0) ...), fenepmp_jbp.py:519 with lambda_D != 0)
        if isinstance(self.a, Coef):
            h = ctx.coef_hess(self.a.fn, self.a.comps)  # (ne,nq,ncomp,2,2)
            return h[..., 0, :, :] if self.a.scalar else h
        raise NotImplementedError("gradient of Grad(%s)" % type(self.a).__name__)

    def deg(self):
        return _dred(self.a.deg())

    def split(self, num):
        return _linear_split(Grad.make, self.a, num)


class NablaGrad(Expr):
    """nabla_grad(f)[k, ...] = d f[...] / d x_k  (derivative index first)."""

    def __init__(self, a):
        self.a = a
        self.shape = (2,) + a.shape
        self.deps = a.deps

    @staticmethod
    def make(a):
        if _is_zero(a) or isinstance(a, Const):
            return Zero((2,) + a.shape)
        return NablaGrad(a)

    def _ev(self, ctx):
        return np.moveaxis(self.a.gr(ctx), -1, 2)

    def deg(self):
        return _dred(self.a.deg())

    def split(self, num):
        return _linear_split(NablaGrad.make, self.a, num)


class DivOp(Expr):
    """div(f)[...] = sum_k d f[..., k] / d x_k  (contracts the LAST index)."""

    def __init__(self, a):
        self.a = a
        self.shape = a.shape[:-1]
        self.deps = a.deps

    @staticmethod
    def make(a):
        if not a.shape or a.shape[-1] != 2:
            raise ValueError("div needs last dimension 2")
        if _is_zero(a) or isinstance(a, Const):
            return Zero(a.shape[:-1])
        return DivOp(a)

    def _ev(self, ctx):
        g = self.a.gr(ctx)  # (e,q,*S,2[last of a],2[deriv])
        return g[..., 0, 0] + g[..., 1, 1]

    def deg(self):
        return _dred(self.a.deg())

    def split(self, num):
        return _linear_split(DivOp.make, self.a, num)


class Dx(Expr):
    def __init__(self, a, i):
        self.a, self.i = a, i
        self.shape = a.shape
        self.deps = a.deps

    @staticmethod
    def make(a, i):
        if _is_zero(a) or isinstance(a, Const):
            return Zero(a.shape)
        return Dx(a, i)

    def _ev(self, ctx):
        return self.a.gr(ctx)[..., self.i]

    def deg(self):
        return _dred(self.a.deg())

    def split(self, num):
        return _linear_split(lambda x: Dx.make(x, self.i), self.a, num)


def grad(a):
    return Grad.make(as_expr(a))


def nabla_grad(a):
    return NablaGrad.make(as_expr(a))


def div(a):
    return DivOp.make(as_expr(a))


def tr(a):
    return a[0, 0] + a[1, 1]


# ---- forms -------------------------------------------------------------------------------------
class Measure:
    def __init__(self, kind, subdomain_id=None):
        self.kind, self.subdomain_id = kind, subdomain_id

    def __call__(self, i):
        return Measure(self.kind, int(i))

    def __getitem__(self, markers):  # Measure("ds")[boundary_parts] (legacy dolfin idiom)
        return self

    def __rmul__(self, e):
        return Form([Integral(as_expr(e), self.kind, self.subdomain_id)])


dx = Measure("dx")
ds = Measure("ds")


class Integral:
    def __init__(self, integrand, kind, subdomain_id):
        if integrand.shape != ():
            raise ValueError("integrand must be scalar")
        self.integrand, self.kind, self.subdomain_id = integrand, kind, subdomain_id


class Form:
    def __init__(self, integrals):
        self.integrals = [i for i in integrals if not _is_zero(i.integrand)]

    def __add__(self, o):
        if isinstance(o, (int, float)) and o == 0:
            return self
        return Form(self.integrals + o.integrals)

    __radd__ = __add__

    def __neg__(self):
        return Form([Integral(-i.integrand, i.kind, i.subdomain_id) for i in self.integrals])

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, s):
        s = as_expr(s)
        return Form([Integral(s * i.integrand, i.kind, i.subdomain_id) for i in self.integrals])

    __rmul__ = __mul__

    def arguments(self):
        d = frozenset()
        for i in self.integrals:
            d |= i.integrand.deps
        return d

    def rank(self):
        return len(self.arguments())


def lhs(F):
    """Bilinear part of a residual form (UFL lhs): integrand terms linear in the trial function."""
    out = []
    for i in F.integrals:
        _, p1 = i.integrand.split(1)
        if p1 is not None:
            out.append(Integral(p1, i.kind, i.subdomain_id))
    return Form(out)


def rhs(F):
    """MINUS the trial-free part of a residual form (UFL rhs)."""
    out = []
    for i in F.integrals:
        p0, _ = i.integrand.split(1)
        if p0 is not None and 0 in p0.deps:
            out.append(Integral(-p0, i.kind, i.subdomain_id))
    return Form(out)
