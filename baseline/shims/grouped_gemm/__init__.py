"""Stand-in for the third-party ``nv-grouped-gemm`` wheel (not installable offline) - REFERENCE ARM ONLY.

Exposes the one symbol the reference uses (``grouped_gemm.backend.gmm``) on top of PyTorch's own library grouped GEMM
(``torch._grouped_mm``, CUTLASS inside libtorch), with a per-expert cuBLAS loop as the fallback.  Nothing here comes
from ``d9d_b200``.
"""

from . import backend  # noqa: F401
