"""vilma_b200 -- B200-native implementation of the `vilma fit` hot path.

Drop-in for ``vilma.variational_inference.MultiPopVI`` over
``vilma.matrix_structures.BlockDiagonalMatrix`` (jeffspence/vilma v0.0.16): Python host
code mirrors the reference interface; every array operation of the fitting loop runs in
hand-written sm_100a CUDA kernels reached through the C ABI in ``include/vilma_b200.h``
(ctypes).  There is no CPU fallback: constructing a fit without the CUDA library or a
device raises.
"""
VERSION = '0.1.0'
REFERENCE_VERSION = '0.0.16'
