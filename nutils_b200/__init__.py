'''B200-native element-integration engine behind the Nutils integrate/integral API.

Scope: the Sample.integrate / Topology.integral -> per-element quadrature ->
sparse (CSR) assembly hot path of evalf/nutils, re-designed for NVIDIA B200
(sm_100a).  Host code is Python; the compute path is hand-written CUDA behind a
ctypes C-ABI (include/b200fem.h, nutils_b200/csrc).  There is no CPU fallback:
operations that need the CUDA library raise if it is missing.
'''

__version__ = '0.1'
