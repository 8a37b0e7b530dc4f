// d9d_b200 — MoE router: fp32 softmax over the experts + top-k selection (+ optional selection bias, + optional
// renormalisation of the selected probabilities) in one pass, one warp per token; and its backward.
// Replaces softmax / topk (radix select + sort) / gather / div / sum launches of the eager path.
#include <stdexcept>

#include "common.cuh"
#include "d9d_ops.h"

namespace d9d {
namespace {

constexpr float RENORM_EPS = 1e-20f;

template <int VPL>  // experts per lane: E <= 32 * VPL
__device__ __forceinline__ void load_row_softmax(const __nv_bfloat16* __restrict__ row, int E, int lane, float (&p)[VPL]) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e = lane + i * 32;  // strided ownership keeps global loads coalesced
    p[i] = (e < E) ? __bfloat162float(row[e]) : -INFINITY;
    mx = fmaxf(mx, p[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    p[i] = (lane + i * 32 < E) ? __expf(p[i] - mx) : 0.f;
    sum += p[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < VPL; ++i) p[i] *= inv;
}

template <int VPL>
__global__ void __launch_bounds__(256) router_fwd_kernel(const __nv_bfloat16* __restrict__ logits,
                                                         const float* __restrict__ bias, long long T, int E, int k,
                                                         bool renorm, long long* __restrict__ idx_out,
                                                         float* __restrict__ probs_out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long t = warp; t < T; t += warps) {
    float p[VPL], key[VPL];
    load_row_softmax<VPL>(logits + t * E, E, lane, p);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int e = lane + i * 32;
      key[i] = (e < E) ? (bias ? p[i] + bias[e] : p[i]) : -INFINITY;
    }
    float sel_p = 0.f;   // lane j keeps the j-th selection
    int sel_e = 0;
    float total = 0.f;
    for (int j = 0; j < k; ++j) {
      // arg-max over the warp; ties go to the lower expert index
      float best = -INFINITY;
      int best_e = 0x7fffffff;
      float best_p = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int e = lane + i * 32;
        if (key[i] > best || (key[i] == best && e < best_e)) { best = key[i]; best_e = e; best_p = p[i]; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oe = __shfl_xor_sync(0xffffffffu, best_e, off);
        const float op = __shfl_xor_sync(0xffffffffu, best_p, off);
        if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; best_p = op; }
      }
      if ((best_e & 31) == lane) key[best_e >> 5] = -INFINITY;  // owner retires the winner
      if (lane == j) { sel_p = best_p; sel_e = best_e; }
      total += best_p;
    }
    if (lane < k) {
      idx_out[t * k + lane] = sel_e;
      probs_out[t * k + lane] = renorm ? sel_p / (total + RENORM_EPS) : sel_p;
    }
  }
}

// dlogits[e] = p_e * (dp_e - sum_j ds_j s_j), dp = scatter(ds), ds_j = (dr_j - sum_i dr_i r_i) / (S + eps) when renorm
template <int VPL>
__global__ void __launch_bounds__(256) router_bwd_kernel(const __nv_bfloat16* __restrict__ logits,
                                                         const long long* __restrict__ idx,
                                                         const float* __restrict__ dprobs, long long T, int E, int k,
                                                         bool renorm, __nv_bfloat16* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long t = warp; t < T; t += warps) {
    float p[VPL];
    load_row_softmax<VPL>(logits + t * E, E, lane, p);
    // lane j < k handles selection j
    const int my_e = (lane < k) ? static_cast<int>(idx[t * k + lane]) : -1;
    float my_dr = (lane < k) ? dprobs[t * k + lane] : 0.f;
    // s_j = p[my_e]: fetch from the owning lane
    float my_s = 0.f;
    for (int j = 0; j < k; ++j) {
      const int e = __shfl_sync(0xffffffffu, my_e, j);
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if ((e >> 5) == i) v = p[i];
      v = __shfl_sync(0xffffffffu, v, e & 31);
      if (lane == j) my_s = v;
    }
    float my_ds = my_dr;
    if (renorm) {
      const float S = warp_sum(my_s) + RENORM_EPS;
      const float r = my_s / S;
      const float dot = warp_sum(my_dr * r);
      my_ds = (lane < k) ? (my_dr - dot) / S : 0.f;
    }
    const float inner = warp_sum(my_ds * my_s);
    float dp[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) dp[i] = 0.f;
    for (int j = 0; j < k; ++j) {
      const int e = __shfl_sync(0xffffffffu, my_e, j);
      const float d = __shfl_sync(0xffffffffu, my_ds, j);
      if ((e & 31) == lane) {
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if ((e >> 5) == i) dp[i] += d;
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int e = lane + i * 32;
      if (e < E) dlogits[t * E + e] = __float2bfloat16_rn(p[i] * (dp[i] - inner));
    }
  }
}

inline int sms() {
  static int n = 0;
  if (!n) { int d; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); }
  return n;
}

inline int grid_for(long long T) {
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(sms()) * 8;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

#define D9D_ROUTER_DISPATCH(KERNEL, ...)                                                     \
  do {                                                                                       \
    if (E <= 32) KERNEL<1><<<grid_for(T), 256, 0, stream>>>(__VA_ARGS__);                    \
    else if (E <= 64) KERNEL<2><<<grid_for(T), 256, 0, stream>>>(__VA_ARGS__);               \
    else if (E <= 128) KERNEL<4><<<grid_for(T), 256, 0, stream>>>(__VA_ARGS__);              \
    else if (E <= 256) KERNEL<8><<<grid_for(T), 256, 0, stream>>>(__VA_ARGS__);              \
    else if (E <= 512) KERNEL<16><<<grid_for(T), 256, 0, stream>>>(__VA_ARGS__);             \
    else KERNEL<32><<<grid_for(T), 256, 0, stream>>>(__VA_ARGS__);                           \
  } while (0)

void router_topk_fwd(const void* logits, const float* bias, long long T, int E, int k, bool renorm, long long* idx,
                     float* probs, cudaStream_t stream) {
  if (T == 0) return;
  if (E > 1024 || k > 32 || k > E) throw std::runtime_error("d9d router: E <= 1024 and k <= min(32, E) required");
  D9D_ROUTER_DISPATCH(router_fwd_kernel, static_cast<const __nv_bfloat16*>(logits), bias, T, E, k, renorm, idx, probs);
}

void router_topk_bwd(const void* logits, const long long* idx, const float* dprobs, long long T, int E, int k,
                     bool renorm, void* dlogits, cudaStream_t stream) {
  if (T == 0) return;
  if (E > 1024 || k > 32 || k > E) throw std::runtime_error("d9d router: E <= 1024 and k <= min(32, E) required");
  D9D_ROUTER_DISPATCH(router_bwd_kernel, static_cast<const __nv_bfloat16*>(logits), idx, dprobs, T, E, k, renorm,
                      static_cast<__nv_bfloat16*>(dlogits));
}

}  // namespace d9d
