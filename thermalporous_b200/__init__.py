"""thermalporous_b200 - B200-native hot path of tlroy/thermalporous.

The per-Newton-step DG0/TPFA residual + Jacobian assembly and the CPR/CPTR-preconditioned
(F)GMRES solve run as hand-written sm_100a CUDA kernels inside libtpb200.so (C ABI in
include/tpb200.h); this package is the host-side mirror of the reference's model / geo / case /
option surface.  There is no CPU fallback.
"""
__version__ = "0.1.0"
