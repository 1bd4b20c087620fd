"""Builds tests/hostemu/libfx_hostemu.so (CPU emulation of the CUDA element kernels -- test
infrastructure only, never imported by fenics_b200)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libfx_hostemu.so")
SRC = os.path.join(HERE, "hostemu.cpp")
CSRC = os.path.join(HERE, "..", "..", "fenics_b200", "csrc")


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    deps.append(os.path.join(HERE, "..", "..", "include", "fenics_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if force or _stale():
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", SO, SRC, "-lpthread"])
    return SO


if __name__ == "__main__":
    print(build(force=True))
