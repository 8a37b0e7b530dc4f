// d9d_b200 — MoE routing layout + token permutation, fully device-side (no host sync, CUDA-graph friendly).
//
// Layout: rows are sorted by expert; every expert segment starts at a multiple of `align` (128 = GEMM BLOCK_M)
// so that a GEMM M-tile never straddles two experts and per-expert K-ranges (wgrad) are BLOCK_K aligned.
// Pad rows are zero.  The sort is a *stable* counting sort (deterministic row order => deterministic wgrad sums).
#include <stdexcept>

#include "common.cuh"
#include "d9d_ops.h"

namespace d9d {
namespace {

constexpr int CHUNK = 1024;  // entries per warp-chunk (32 rounds of 32)

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// pass 1: per-chunk histogram
__global__ void __launch_bounds__(256) moe_hist_kernel(const long long* __restrict__ ids, long long n, int E,
                                                       int* __restrict__ chunk_hist) {
  extern __shared__ int sh[];  // [warps_per_block][E]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long chunk = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  int* my = sh + warp * E;
  for (int e = lane; e < E; e += 32) my[e] = 0;
  __syncwarp();
  const long long begin = chunk * CHUNK;
  const long long end = min(begin + CHUNK, n);
  for (long long i = begin + lane; i < end; i += 32) {
    const long long e = ids[i];
    if (e >= 0 && e < E) atomicAdd(&my[e], 1);
  }
  __syncwarp();
  if (begin < n)
    for (int e = lane; e < E; e += 32) chunk_hist[chunk * E + e] = my[e];
}

// pass 2 (single block): scan chunks per expert, aligned segment offsets, tile->expert table
__global__ void __launch_bounds__(1024) moe_scan_kernel(int* __restrict__ chunk_hist, long long num_chunks, int E,
                                                        int align, long long capacity_rows, int* __restrict__ counts,
                                                        int* __restrict__ seg_offsets, int* __restrict__ tile_group) {
  __shared__ int s_count[1024];
  __shared__ int s_off[1025];
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int run = 0;
    for (long long c = 0; c < num_chunks; ++c) {
      const int h = chunk_hist[c * E + e];
      chunk_hist[c * E + e] = run;  // exclusive prefix within the expert
      run += h;
    }
    s_count[e] = run;
    counts[e] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int e = 0; e < E; ++e) {
      s_off[e] = off;
      off += (s_count[e] + align - 1) / align * align;
    }
    s_off[E] = off;
  }
  __syncthreads();
  for (int e = threadIdx.x; e <= E; e += blockDim.x) seg_offsets[e] = s_off[e];
  // the tile->expert table only exists for GEMM-aligned layouts (compact align=1 layouts feed the all-to-all)
  const int num_tiles = (align % 128 == 0) ? static_cast<int>(capacity_rows / 128) : 0;
  for (int t = threadIdx.x; t < num_tiles; t += blockDim.x) tile_group[t] = -1;
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int t0 = s_off[e] / 128, t1 = s_off[e + 1] / 128;
    for (int t = t0; t < t1 && t < num_tiles; ++t) tile_group[t] = e;
  }
}

// pass 3: stable destination rows
__global__ void __launch_bounds__(256) moe_rowmap_kernel(const long long* __restrict__ ids, long long n, int E,
                                                         const int* __restrict__ chunk_prefix,
                                                         const int* __restrict__ seg_offsets,
                                                         int* __restrict__ row_map) {
  extern __shared__ int sh[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long chunk = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  int* run = sh + warp * E;
  const long long begin = chunk * CHUNK;
  if (begin >= n) return;
  for (int e = lane; e < E; e += 32) run[e] = seg_offsets[e] + chunk_prefix[chunk * E + e];
  __syncwarp();
  const long long end = min(begin + CHUNK, n);
  for (long long base = begin; base < end; base += 32) {
    const long long i = base + lane;
    const bool in = i < end;
    const long long e64 = in ? ids[i] : -1;
    const bool valid = in && e64 >= 0 && e64 < E;
    const int e = valid ? static_cast<int>(e64) : -1 - lane;  // unique negative keys never match each other
    const uint32_t peers = __match_any_sync(0xffffffffu, e);
    const int rank = __popc(peers & lanemask_lt());
    int dst = -1;
    if (valid) dst = run[e] + rank;
    __syncwarp();
    if (valid && rank == __popc(peers) - 1) run[e] += __popc(peers);  // last peer bumps the running cursor
    __syncwarp();
    if (in) row_map[i] = dst;
  }
}

// one warp per token: read the row once, scatter to its k destinations
__global__ void __launch_bounds__(256) moe_scatter_kernel(const __nv_bfloat16* __restrict__ x,
                                                          const float* __restrict__ probs,
                                                          const int* __restrict__ row_map,
                                                          __nv_bfloat16* __restrict__ xp, float* __restrict__ pp,
                                                          long long T, int k, int H) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long warps_total = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = H >> 3;
  for (long long t = warp_global; t < T; t += warps_total) {
    const uint4* src = reinterpret_cast<const uint4*>(x + t * H);
    for (int v0 = 0; v0 < nvec; v0 += 32 * 4) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * 32 + lane;
        if (v < nvec) r[u] = src[v];
      }
      for (int j = 0; j < k; ++j) {
        const int dst = row_map[t * k + j];
        if (dst < 0) continue;
        uint4* d = reinterpret_cast<uint4*>(xp + static_cast<long long>(dst) * H);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = v0 + u * 32 + lane;
          if (v < nvec) d[v] = r[u];
        }
      }
    }
    if (pp != nullptr && lane < k) {
      const int dst = row_map[t * k + lane];
      if (dst >= 0) pp[dst] = probs[t * k + lane];
    }
  }
}

// one warp per token: y[t] = sum_j yp[row_map[t,j]] (fp32 accumulate); optional dprobs gather
__global__ void __launch_bounds__(256) moe_gather_sum_kernel(const __nv_bfloat16* __restrict__ yp,
                                                             const float* __restrict__ dpp,
                                                             const int* __restrict__ row_map,
                                                             __nv_bfloat16* __restrict__ y,
                                                             float* __restrict__ dprobs, long long T, int k, int H) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long warps_total = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = H >> 3;
  for (long long t = warp_global; t < T; t += warps_total) {
    for (int v = lane; v < nvec; v += 32) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int j = 0; j < k; ++j) {
        const int src = row_map[t * k + j];
        if (src < 0) continue;
        const uint4 u = reinterpret_cast<const uint4*>(yp + static_cast<long long>(src) * H)[v];
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      }
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
      reinterpret_cast<uint4*>(y + t * H)[v] = o;
    }
    if (dprobs != nullptr && lane < k) {
      const int src = row_map[t * k + lane];
      dprobs[t * k + lane] = (src >= 0) ? dpp[src] : 0.f;
    }
  }
}

// zero the pad rows of every expert segment (one block per expert)
__global__ void __launch_bounds__(256) moe_zero_pad_kernel(__nv_bfloat16* __restrict__ xp, float* __restrict__ pp,
                                                           const int* __restrict__ counts,
                                                           const int* __restrict__ seg_offsets, int H) {
  const int e = blockIdx.x;
  const int r0 = seg_offsets[e] + counts[e], r1 = seg_offsets[e + 1];
  const int nvec = H >> 3;
  const long long total = static_cast<long long>(r1 - r0) * nvec;
  uint4* base = reinterpret_cast<uint4*>(xp + static_cast<long long>(r0) * H);
  for (long long i = threadIdx.x; i < total; i += blockDim.x) base[i] = make_uint4(0, 0, 0, 0);
  if (pp != nullptr)
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) pp[r] = 0.f;
}

inline int num_sms() {
  static int n = 0;
  if (!n) { int d; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); }
  return n;
}

}  // namespace

long long moe_layout_scratch_ints(long long n_entries, int E) {
  const long long chunks = (n_entries + CHUNK - 1) / CHUNK;
  return chunks * E;
}

void moe_build_layout(const long long* topk_ids, long long T, int k, int E, int align, long long capacity_rows,
                      int* counts, int* seg_offsets, int* row_map, int* tile_group, int* chunk_scratch,
                      cudaStream_t stream) {
  if (E > 1024) throw std::runtime_error("d9d moe: at most 1024 local experts supported");
  if (align < 1 || (align % 128 == 0 && capacity_rows % 128 != 0))
    throw std::runtime_error("d9d moe: capacity of a 128-aligned layout must be a multiple of 128");
  const long long n = T * k;
  const long long chunks = (n + CHUNK - 1) / CHUNK;
  const int warps_per_block = 8;
  const int blocks = static_cast<int>((chunks + warps_per_block - 1) / warps_per_block);
  const size_t smem = static_cast<size_t>(warps_per_block) * E * sizeof(int);
  if (blocks > 0) {
    moe_hist_kernel<<<blocks, 256, smem, stream>>>(topk_ids, n, E, chunk_scratch);
  }
  moe_scan_kernel<<<1, 1024, 0, stream>>>(chunk_scratch, chunks, E, align, capacity_rows, counts, seg_offsets,
                                          tile_group);
  if (blocks > 0) {
    moe_rowmap_kernel<<<blocks, 256, smem, stream>>>(topk_ids, n, E, chunk_scratch, seg_offsets, row_map);
  }
}

void moe_permute(const void* x, const float* probs, const int* row_map, void* xp, float* pp, long long T, int k,
                 int H, cudaStream_t stream) {
  if (T == 0) return;
  if (H % 8 != 0 || k > 32) throw std::runtime_error("d9d moe_permute: H % 8 == 0 and k <= 32 required");
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  moe_scatter_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), probs, row_map, static_cast<__nv_bfloat16*>(xp), pp, T, k, H);
}

void moe_zero_pad(void* xp, float* pp, const int* counts, const int* seg_offsets, int E, int H, cudaStream_t stream) {
  if (E == 0) return;
  moe_zero_pad_kernel<<<E, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(xp), pp, counts, seg_offsets, H);
}

void moe_unpermute(const void* yp, const int* row_map, void* y, long long T, int k, int H, cudaStream_t stream) {
  if (T == 0) return;
  if (H % 8 != 0 || k > 32) throw std::runtime_error("d9d moe_unpermute: H % 8 == 0 and k <= 32 required");
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  moe_gather_sum_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(yp), nullptr, row_map, static_cast<__nv_bfloat16*>(y), nullptr, T, k, H);
}

void moe_permute_bwd(const void* dxp, const float* dpp, const int* row_map, void* dx, float* dprobs, long long T,
                     int k, int H, cudaStream_t stream) {
  if (T == 0) return;
  if (H % 8 != 0 || k > 32) throw std::runtime_error("d9d moe_permute_bwd: H % 8 == 0 and k <= 32 required");
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  moe_gather_sum_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dxp), dpp, row_map, static_cast<__nv_bfloat16*>(dx), dprobs, T, k, H);
}

}  // namespace d9d
