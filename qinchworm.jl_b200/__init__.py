"""qinchworm_b200 — B200-native qMC diagram-evaluation hot path of QInchworm.jl.

Host-side mirror (Python, because Julia is not in this image) of the reference's interface for
the hot path, above the C-ABI library `libqinchworm_cuda.so` (csrc/, include/qinchworm.h).
The compute path is CUDA only: there is no CPU fallback; importing `lib` without the built
library raises.
"""
__version__ = "0.1.0"
