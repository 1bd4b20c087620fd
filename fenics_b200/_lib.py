"""ctypes binding of libfenics_b200.so (the C ABI in include/fenics_b200.h).

There is no CPU fallback: importing this module without the built CUDA library raises.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# FX_LIB: another build of the same library (tuning experiments: tools/build_variants.sh), never a different backend
SO_PATH = os.environ.get("FX_LIB") or os.path.join(_HERE, "libfenics_b200.so")


class FxError(RuntimeError):
    pass


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("residual", C.c_double),
                ("residual0", C.c_double)]


def _load():
    if not os.path.exists(SO_PATH):
        raise FxError("libfenics_b200.so is not built (run `python -m fenics_b200.build`); "
                      "fenics_b200 has no CPU fallback")
    lib = C.CDLL(SO_PATH)
    vp, ip, dp, i, d = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_double
    sig = {
        "fx_last_error": (C.c_char_p, []),
        "fx_version": (i, []),
        "fx_launch_count": (C.c_longlong, []),
        "fx_copy_to_host": (i, [vp, vp, C.c_longlong]),
        "fx_plan_create": (i, [vp, i, vp, i, vp, vp, vp, vp, i, i, C.POINTER(vp)]),
        "fx_plan_destroy": (None, [vp]),
        "fx_plan_set_markers": (i, [vp, vp, i]),
        "fx_space_dim": (i, [vp, i]),
        "fx_space_local_dim": (i, [i]),
        "fx_pattern": (i, [vp, i, ip, ip, C.POINTER(vp), C.POINTER(vp)]),
        "fx_geometry": (i, [vp, C.POINTER(vp)]),
        "fx_assemble_matrix": (i, [vp, i, vp, vp, vp, vp]),
        "fx_assemble_vector": (i, [vp, i, vp, vp, vp, vp]),
        "fx_assemble_scalar": (i, [vp, i, vp, vp, dp, vp]),
        "fx_assemble_scalar_dev": (i, [vp, i, vp, vp, vp, vp]),
        "fx_project_dg0": (i, [vp, i, vp, vp, vp, vp]),
        "fx_form_info": (i, [i, ip, ip, ip]),
        "fx_apply_bc": (i, [vp, i, vp, vp, vp, vp, i, vp]),
        "fx_spmv": (i, [vp, i, vp, vp, vp, vp]),
        "fx_spmv_variant": (i, [vp, i, vp, vp, vp, i, vp]),
        "fx_axpy": (i, [i, d, vp, vp, vp]),
        "fx_dot": (i, [vp, i, vp, vp, dp, vp]),
        "fx_norm": (i, [vp, i, vp, i, dp, vp]),
        "fx_dot_dev": (i, [vp, i, vp, vp, vp, vp]),
        "fx_norm_dev": (i, [vp, i, vp, i, vp, vp]),
        "fx_lincomb3_dev": (i, [i, vp, vp, vp, vp, vp, vp]),
        "fx_dotw_dev": (i, [vp, i, vp, vp, vp, vp, vp]),
        "fx_plan_set_owned_cells": (i, [vp, i]),
        "fx_plan_set_assembly_mode": (i, [vp, i]),
        "fx_space_solver_index": (i, [vp, i, vp, ip]),
        "fx_plan_set_halo_solver_index": (i, [vp, i, i, vp]),
        "fx_plan_set_eliminated_dofs": (i, [vp, i, vp, i]),
        "fx_plan_set_aggregates": (i, [vp, i, vp, i]),
        "fx_peer_region_bytes": (C.c_longlong, [C.c_longlong]),
        "fx_plan_set_peer_regions": (i, [vp, i, i, vp, C.c_longlong]),
        "fx_plan_set_halo": (i, [vp, i, i, vp, vp, vp, vp]),
        "fx_krylov_solve": (i, [vp, i, vp, vp, vp, i, i, d, d, i, C.POINTER(SolveInfo), vp]),
        "fx_krylov_solve_async": (i, [vp, i, vp, vp, vp, i, i, d, d, i, i, vp]),
        "fx_solve_results": (i, [vp, i, i, C.POINTER(SolveInfo), vp]),
        "fx_solve_results_async": (i, [vp, i, i, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    lib._fx_names = sorted(sig)
    return lib


lib = _load()
EXPORTS = lib._fx_names


def check(rc):
    if rc != 0:
        msg = lib.fx_last_error().decode()
        if rc == _abi.FX_ERR_NOCONV:
            raise RuntimeError("*** Error: Unable to solve linear system (Krylov solver did not converge): " + msg)
        raise FxError("fenics_b200 error %d: %s" % (rc, msg))
