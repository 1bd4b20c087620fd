"""fenics_b200 -- B200-native drop-in for the per-time-step assembly + linear-solve hot path of the
ATMackay/fenics viscoelastic-flow scripts.  Hand-written sm_100a CUDA kernels behind
the C ABI of include/fenics_b200.h; there is no CPU fallback."""
__version__ = "0.1.0"
