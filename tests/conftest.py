import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # test infrastructure only: a fresh checkout has no built library yet (the product path itself never builds or
    # falls back -- fenics_b200._lib raises when libfenics_b200.so is missing)
    so = os.path.join(ROOT, "fenics_b200", "libfenics_b200.so")
    if not os.path.exists(so):
        import importlib.util
        spec = importlib.util.spec_from_file_location("fx_build", os.path.join(ROOT, "fenics_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
