"""B200-native (sm_100a, FP64) implementation of the NMGP DSVI hot path.

Drop-in for the hot path of Corleno/Collaborative_Nonstationary_Multivariate_Gaussian_Process:
the module names mirror the reference (``utils``, ``nmgp_dsvi``; SIM_code ``kernels``,
``kronecker_operation``, ``distributions``); all arithmetic runs in hand-written CUDA
kernels behind the C ABI of ``include/nmgp_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
