"""B200-native (sm_100a) implementation of LaDDer's ELBO forward/backward hot path.

Layout: `csrc/` CUDA kernels + C ABI (`libladder_sm100.so`, declared in
`include/ladder_sm100.h`), `lib.py` ctypes binding, `ops.py` tensor-level wrappers,
`host/` the mirror of the reference's `codes/` interface (models, trainers, utils).
"""
__version__ = '0.1.0'
