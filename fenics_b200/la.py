"""Device-resident plan, vectors and matrices behind the dolfin-like API (PyTorch tensors are used as
device buffers and for stream handles only; all arithmetic is in libfenics_b200.so)."""
import ctypes as C

import numpy as np
import torch

from . import _abi as abi
from ._lib import SolveInfo, check, lib

_KSP = {"bicgstab": abi.FX_KSP_BICGSTAB, "cg": abi.FX_KSP_CG, "gmres": abi.FX_KSP_GMRES}
_PC = {"none": abi.FX_PC_NONE, "jacobi": abi.FX_PC_JACOBI, "ilu": abi.FX_PC_ILU0, "jacobi_coarse": abi.FX_PC_JACOBI_COARSE}
# dolfin preconditioner names that are SUBSTITUTED, never silently: the first use of each name warns
# (`parameters["preconditioner_substitution"]`: "warn" (default) | "silent" | "error").
#   "default" (PETSc, serial: ILU(0))  -> "jacobi"        e.g. solve(A, x, b, "bicgstab", "default"), incomp_LDC.py:331
#                                         or, with parameters["default_preconditioner"] = "ilu", the library's own
#                                         natural-ordering ILU(0) (csrc/fx_ilu.cu): PETSc's operator and iteration
#                                         counts, at the price of its serial dependency chain
#   "amg" / "petsc_amg" / "hypre_amg"  -> "jacobi_coarse" for CG on a space with aggregates (fx_plan_set_aggregates),
#                                         else "jacobi"   e.g. KrylovSolver("cg", "amg"), fenepmp_jbp.py:437-445
# With the substitutes iteration counts differ from PETSc's (ILU(0) needs 3-5x fewer iterations than Jacobi on the W
# systems); the stopping rule (left-preconditioned residual, rtol 1e-5) is the same.
_PC_SUBSTITUTE = {"default": "jacobi", "amg": "jacobi_coarse", "petsc_amg": "jacobi_coarse", "hypre_amg": "jacobi_coarse"}
_warned = set()


class PreconditionerSubstitution(UserWarning):
    pass


def resolve_pc(pc, method, A):
    """name -> name of a preconditioner the library implements (see _PC_SUBSTITUTE)."""
    if pc in _PC:
        return pc
    if pc == "default" and parameters.get("default_preconditioner", "jacobi") == "ilu":
        return "ilu"
    if pc not in _PC_SUBSTITUTE:
        raise ValueError("unknown method/preconditioner %r/%r" % (method, pc))
    to = _PC_SUBSTITUTE[pc]
    if to == "jacobi_coarse" and not (method == "cg" and A.plan.has_aggregates(A.space)):
        to = "jacobi"
    mode = parameters.get("preconditioner_substitution", "warn")
    if mode == "error":
        raise ValueError("preconditioner %r is not implemented (would be served by %r)" % (pc, to))
    if mode == "warn" and (pc, to) not in _warned:
        import warnings
        _warned.add((pc, to))
        warnings.warn("fenics_b200: preconditioner %r is served by %r; iteration counts differ from PETSc's%s"
                      % (pc, to, " (parameters['default_preconditioner'] = 'ilu' selects the ILU(0) kernels)"
                         if pc == "default" else " (no AMG kernel)"), PreconditionerSubstitution, stacklevel=3)
    return to


# mirror of dolfin's global `parameters["krylov_solver"]` (PETSc defaults)
parameters = {"krylov_solver": {"relative_tolerance": 1e-5, "absolute_tolerance": 1e-50,
                                "maximum_iterations": 10000, "nonzero_initial_guess": True,
                                "monitor_convergence": False, "error_on_nonconvergence": True},
              "std_out_all_processes": False, "preconditioner_substitution": "warn", "default_preconditioner": "jacobi"}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Plan:
    """fx_plan wrapper: mesh + dofmaps on the device (one per mesh; replaces what dolfin rebuilds on
    every assemble() call)."""

    def __init__(self, mesh, spaces, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fenics_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.mesh = mesh
        self.spaces = {s.name: s for s in spaces if s is not None}
        dm = (C.c_void_p * abi.NSPACES)()
        self._keep = []
        for name, s in self.spaces.items():
            a = np.ascontiguousarray(s.cell_dofs, dtype=np.int32)
            assert a.shape == (mesh.num_cells(), abi.SPACE_LD[name])
            self._keep.append(a)
            dm[abi.SPACE_ID[name]] = a.ctypes.data
        xy = np.ascontiguousarray(mesh.coordinates(), dtype=np.float64)
        cells = np.ascontiguousarray(mesh.cells(), dtype=np.int32)
        bc_, bl, bm = (np.ascontiguousarray(a, dtype=np.int32) for a in
                       (mesh.bfacet_cell, mesh.bfacet_local, mesh.bfacet_marker))
        h = C.c_void_p()
        check(lib.fx_plan_create(xy.ctypes.data, len(xy), cells.ctypes.data, len(cells), dm, bc_.ctypes.data,
                                 bl.ctypes.data, bm.ctypes.data, len(bc_), self.device, C.byref(h)))
        self.h = h
        self._pat = {}
        self._agg = set()
        for s in self.spaces.values():  # TestFunction(V) / literal forms find the device plan through their space
            s._shim_plan = self

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h and lib is not None:  # `lib` is already None during interpreter shutdown
            lib.fx_plan_destroy(h)

    def set_markers(self, markers):
        m = np.ascontiguousarray(markers, dtype=np.int32)
        check(lib.fx_plan_set_markers(self.h, m.ctypes.data, len(m)))

    def set_assembly_mode(self, mode):
        """"atomic" (default: fp64 atomicAdd scatter) or "deterministic" (atomic-free row gather, bitwise
        reproducible): how assemble() adds element tensors into the global tensor (fx_plan_set_assembly_mode)."""
        m = {"atomic": abi.FX_ASM_ATOMIC, "deterministic": abi.FX_ASM_DETERMINISTIC}[mode]
        check(lib.fx_plan_set_assembly_mode(self.h, m))

    def set_eliminated_dofs(self, space, rows):
        """Krylov solves on `space` iterate on the other rows and back-substitute these (block-triangular
        systems, fx_plan_set_eliminated_dofs); an empty list switches it off."""
        r = np.ascontiguousarray(rows, dtype=np.int32)
        check(lib.fx_plan_set_eliminated_dofs(self.h, abi.SPACE_ID[space], r.ctypes.data, len(r)))

    def set_aggregates(self, space, agg, k):
        """aggregates of the two-level preconditioner "jacobi_coarse" (fx_plan_set_aggregates); returns False
        (and leaves the space without aggregates) when the library cannot use them for this space."""
        a = np.ascontiguousarray(agg, dtype=np.int32)
        rc = lib.fx_plan_set_aggregates(self.h, abi.SPACE_ID[space], a.ctypes.data, int(k))
        self._agg.discard(space)
        if rc == abi.FX_ERR_UNSUPPORTED:
            return False
        check(rc)
        if int(k) > 0:
            self._agg.add(space)
        return True

    def has_aggregates(self, space):
        return space in self._agg

    def dim(self, space):
        return lib.fx_space_dim(self.h, abi.SPACE_ID[space])

    def pattern(self, space):
        """(n, nnz, rowptr_ptr, col_ptr) -- device pointers owned by the plan."""
        if space not in self._pat:
            n, nnz = C.c_int(), C.c_int()
            rp, col = C.c_void_p(), C.c_void_p()
            check(lib.fx_pattern(self.h, abi.SPACE_ID[space], C.byref(n), C.byref(nnz), C.byref(rp), C.byref(col)))
            self._pat[space] = (n.value, nnz.value, rp.value, col.value)
        return self._pat[space]

    def pattern_host(self, space):
        n, nnz, rp, col = self.pattern(space)
        out_rp = np.empty(n + 1, dtype=np.int32)
        out_col = np.empty(nnz, dtype=np.int32)
        check(lib.fx_copy_to_host(out_rp.ctypes.data, rp, 4 * (n + 1)))
        check(lib.fx_copy_to_host(out_col.ctypes.data, col, 4 * nnz))
        return out_rp, out_col

    def geometry_host(self):
        g = C.c_void_p()
        check(lib.fx_geometry(self.h, C.byref(g)))
        nc = self.mesh.num_cells()
        out = np.empty((6, nc))
        check(lib.fx_copy_to_host(out.ctypes.data, g.value, 8 * 6 * nc))
        return out

    def new_vector(self, space):
        return Vector(self, space)

    def new_matrix(self, space):
        return Matrix(self, space)


class Vector:
    """GenericVector stand-in: fp64 device vector in the dof order of `space`."""

    def __init__(self, plan, space, data=None):
        self.plan, self.space = plan, space
        n = plan.dim(space)
        self.t = torch.zeros(n, dtype=torch.float64, device="cuda:%d" % plan.device) if data is None else data

    def size(self):
        return self.t.numel()

    def get_local(self):
        return self.t.cpu().numpy()

    array = get_local

    def set_local(self, a):
        self.t.copy_(torch.as_tensor(np.asarray(a, dtype=np.float64)), non_blocking=False)

    def __setitem__(self, key, value):
        """v[:] = a (the scripts' idiom) or any index / slice / index array torch accepts."""
        if isinstance(key, slice) and key == slice(None):
            self.set_local(value) if np.ndim(value) else self.t.fill_(float(value))
        else:
            if isinstance(key, np.ndarray):
                key = torch.as_tensor(key, device=self.t.device)
            self.t[key] = torch.as_tensor(np.asarray(value, dtype=np.float64), device=self.t.device)

    def zero(self):
        self.t.zero_()

    def copy(self):
        return Vector(self.plan, self.space, self.t.clone())

    def norm(self, kind="l2"):
        out = C.c_double()
        check(lib.fx_norm(self.plan.h, self.size(), _ptr(self.t), 1 if kind == "linf" else 0, C.byref(out), _stream()))
        return out.value

    def inner(self, other):
        out = C.c_double()
        check(lib.fx_dot(self.plan.h, self.size(), _ptr(self.t), _ptr(other.t), C.byref(out), _stream()))
        return out.value

    def axpy(self, alpha, x):
        check(lib.fx_axpy(self.size(), float(alpha), _ptr(x.t), _ptr(self.t), _stream()))


class Matrix:
    """GenericMatrix stand-in: CSR values on the plan's DOLFIN-pattern of space x space."""

    def __init__(self, plan, space):
        self.plan, self.space = plan, space
        n, nnz, _, _ = plan.pattern(space)
        self.n, self.nnz = n, nnz
        self.val = torch.empty(nnz, dtype=torch.float64, device="cuda:%d" % plan.device)

    def mult(self, x, y):
        check(lib.fx_spmv(self.plan.h, abi.SPACE_ID[self.space], _ptr(self.val), _ptr(x.t), _ptr(y.t), _stream()))

    def __mul__(self, x):
        y = Vector(self.plan, self.space)
        self.mult(x, y)
        return y

    def to_scipy(self):
        import scipy.sparse as sp
        rp, col = self.plan.pattern_host(self.space)
        return sp.csr_matrix((self.val.cpu().numpy(), col, rp), shape=(self.n, self.n))

    array = to_scipy


def norm(v, kind="l2"):
    """dolfin norm(vector, 'l2'|'linf') (lid-driven-cavity/incomp_LDC.py:435)."""
    v = v.vector() if hasattr(v, "vector") else v
    return v.norm(kind)


def krylov_solve(A, x, b, method="bicgstab", pc="default", rtol=None, atol=None, maxit=None, ticket=None):
    """solve(A, x, b, method, pc): one persistent cooperative kernel.  ticket=None: synchronous, returns
    SolveInfo and raises RuntimeError like dolfin on non-convergence.  ticket=k: asynchronous, the outcome
    lands in the plan's result slot k (see solve_results)."""
    ks = parameters["krylov_solver"]
    rtol = ks["relative_tolerance"] if rtol is None else rtol
    atol = ks["absolute_tolerance"] if atol is None else atol
    maxit = ks["maximum_iterations"] if maxit is None else maxit
    if not ks["nonzero_initial_guess"]:
        x.t.zero_()
    if method not in _KSP:
        raise ValueError("unknown method/preconditioner %r/%r" % (method, pc))
    pc = resolve_pc(pc, method, A)
    args = (A.plan.h, abi.SPACE_ID[A.space], _ptr(A.val), _ptr(x.t), _ptr(b.t), _KSP[method], _PC[pc],
            float(rtol), float(atol), int(maxit))
    if ticket is not None:
        check(lib.fx_krylov_solve_async(*args, int(ticket), _stream()))
        return None
    info = SolveInfo()
    rc = lib.fx_krylov_solve(*args, C.byref(info), _stream())
    if rc == abi.FX_ERR_NOCONV and not ks["error_on_nonconvergence"]:
        return info
    check(rc)
    return info


def solve_results(plan, first, count):
    """wait for the stream and fetch `count` asynchronous solve outcomes starting at slot `first`."""
    infos = (SolveInfo * count)()
    check(lib.fx_solve_results(plan.h, first, count, infos, _stream()))
    return list(infos)
