"""env_build_b200: B200-native implementation of the CrossroadEnd2end model hot path.

The package mirrors the reference's flat modules for this path
(`dynamics_and_models`, `endtoend_env_utils`, `endtoend`); all arithmetic on
the path runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/ce2e.h (libce2e.so).  Importing the package is cheap; the CUDA library
is loaded on first use and there is NO CPU fallback.
"""
__version__ = '0.1.0'
