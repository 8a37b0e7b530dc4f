"""Stand-in for the third-party ``cut-cross-entropy`` wheel (not installable offline) - REFERENCE ARM ONLY.

Provides the symbols the reference's ``d9d/kernel/cce`` imports.  The two Triton kernels of the real package
(``cce_lse_forward_kernel``, ``cce_backward_kernel``) are replaced by row-chunked PyTorch code on library GEMMs
(cuBLAS): logits are materialised one chunk at a time, gradients are exact (no gradient filtering - the real kernels
skip blocks whose softmax is below ``filter_eps``).  Nothing here comes from ``d9d_b200``.
"""
