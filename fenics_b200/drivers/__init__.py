"""Time-step drivers: the reference scripts' `while t < Tf` loop bodies on the GPU-backed API."""
