"""ctypes wrapper of the host emulation library (tests only)."""
import ctypes as C
import numpy as np

import importlib.util, os
_spec = importlib.util.spec_from_file_location("fx_hostemu_build", os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu", "build.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
build = _mod.build

from fenics_b200 import _abi as _abi_consts  # noqa: E402

NSPACES, NFIELDS, NPARAMS = _abi_consts.NSPACES, _abi_consts.NFIELDS, _abi_consts.NPARAMS
SPACES = {"W": 0, "Zc": 1, "Q": 2, "Zd": 3, "Qt": 4, "V": 5, "Vs": 6}
LD = {"W": 15, "Zc": 18, "Q": 3, "Zd": 3, "Qt": 1, "V": 12, "Vs": 6}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class EmuPlan:
    def __init__(self, coords, cells, dofmaps, bf_cell, bf_local, bf_marker):
        self.lib = C.CDLL(build())
        L = self.lib
        L.emu_plan_create.restype = C.c_void_p
        L.emu_geometry.restype = _dp
        L.emu_facet_geometry.restype = _dp
        self.nc = len(cells)
        self.keep = [np.ascontiguousarray(coords, dtype=np.float64), _i32(cells), _i32(bf_cell), _i32(bf_local),
                     _i32(bf_marker)]
        dm = (C.c_void_p * NSPACES)()
        self.dims = {}
        for name, sid in SPACES.items():
            if name in dofmaps:
                a = _i32(dofmaps[name])
                assert a.shape == (self.nc, LD[name])
                self.keep.append(a)
                dm[sid] = a.ctypes.data
                self.dims[name] = int(a.max()) + 1
        k = self.keep
        self.h = L.emu_plan_create(C.c_void_p(k[0].ctypes.data), len(coords), C.c_void_p(k[1].ctypes.data), self.nc,
                                   dm, C.c_void_p(k[2].ctypes.data), C.c_void_p(k[3].ctypes.data),
                                   C.c_void_p(k[4].ctypes.data), len(bf_cell))
        assert self.h
        self.h = C.c_void_p(self.h)

    def pattern(self, space):
        n, nnz = C.c_int(), C.c_int()
        rp, col = _ip(), _ip()
        self.lib.emu_pattern(self.h, SPACES[space], C.byref(n), C.byref(nnz), C.byref(rp), C.byref(col))
        return (np.ctypeslib.as_array(rp, (n.value + 1,)).copy(), np.ctypeslib.as_array(col, (nnz.value,)).copy())

    def geometry(self):
        return np.ctypeslib.as_array(self.lib.emu_geometry(self.h), (6, self.nc)).copy()

    def _state(self, state):
        arr = (C.c_void_p * NFIELDS)()
        keep = []
        for k, v in state.items():
            v = np.ascontiguousarray(v, dtype=np.float64)
            keep.append(v)
            arr[k] = v.ctypes.data
        return arr, keep

    @staticmethod
    def _params(params):
        p = np.zeros(NPARAMS)
        for k, v in params.items():
            p[k] = v
        return p

    def assemble_matrix(self, form, space, state, params):
        import scipy.sparse as sp
        rp, col = self.pattern(space)
        val = np.zeros(len(col))
        st, keep = self._state(state)
        p = self._params(params)
        rc = self.lib.emu_assemble_matrix(self.h, form, st, C.c_void_p(p.ctypes.data), C.c_void_p(val.ctypes.data))
        assert rc == 0, rc
        return sp.csr_matrix((val, col, rp), shape=(len(rp) - 1,) * 2)

    def assemble_vector(self, form, space, state, params):
        b = np.zeros(self.dims[space])
        st, keep = self._state(state)
        p = self._params(params)
        rc = self.lib.emu_assemble_vector(self.h, form, st, C.c_void_p(p.ctypes.data), C.c_void_p(b.ctypes.data))
        assert rc == 0, rc
        return b

    def assemble_scalar(self, form, state, params):
        out = C.c_double()
        st, keep = self._state(state)
        p = self._params(params)
        rc = self.lib.emu_assemble_scalar(self.h, form, st, C.c_void_p(p.ctypes.data), C.byref(out))
        assert rc == 0, rc
        return out.value

    def project_dg0(self, expr, space, state, params):
        out = np.zeros(self.dims[space])
        st, keep = self._state(state)
        p = self._params(params)
        rc = self.lib.emu_project_dg0(self.h, expr, st, C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data))
        assert rc == 0, rc
        return out

    def form_qdeg(self, form, conv=1.0):
        a, b = C.c_int(), C.c_int()
        rc = self.lib.emu_form_qdeg(self.h if False else form, C.byref(a), C.byref(b), C.c_double(conv))
        assert rc == 0, rc
        return a.value, b.value
