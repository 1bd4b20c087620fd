"""Builds oracle/libfx_oracle.so from oracle/krylov_c.c with gcc (test infrastructure: the CPU restatement of the
reference's PETSc BiCGStab + ILU(0) solve used by tests/ and bench.py's CPU legs)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libfx_oracle.so")
SRC = os.path.join(HERE, "krylov_c.c")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SRC) > os.path.getmtime(SO):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-shared", "-fPIC", "-o", SO, SRC, "-lm"])
    return SO


if __name__ == "__main__":
    print(build(force=True))
