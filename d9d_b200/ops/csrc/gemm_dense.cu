// Dense tcgen05 GEMM instantiations (forward NT, dgrad NN-as-MN-major-B, wgrad TN-as-MN/MN).
#include "gemm_host.cuh"

namespace d9d {
using namespace gemm;

void gemm_dense(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return;
  const long long m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int bn = a.block_n ? a.block_n : pick_block_n(m_tiles, a.N);
  if (!a.a_mn && !a.b_mn) {
    D9D_DISPATCH_EPI4(DENSE, false, false, a.epi, bn, a, stream);
  } else if (!a.a_mn && a.b_mn) {
    D9D_DISPATCH_EPI4(DENSE, false, true, a.epi, bn, a, stream);  // accumulate: chunked dE of the fused linear-CE
  } else if (a.a_mn && a.b_mn) {
    D9D_DISPATCH_EPI4(DENSE, true, true, a.epi, bn, a, stream);
  } else {
    D9D_DISPATCH_EPI2(DENSE, true, false, a.epi, bn, a, stream);
  }
}
}  // namespace d9d
