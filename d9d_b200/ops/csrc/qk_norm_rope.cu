// d9d_b200 — per-head RMSNorm of q and k fused with rotary embedding (forward and backward).
// One warp per (token, head); q and k heads share a launch.  Replaces 2 RMSNorm + 2 RoPE launches (forward) and
// 2 x (RMSNorm backward + dW reduce) + 2 RoPE launches (backward), and halves the activation traffic.
#include <stdexcept>

#include "common.cuh"
#include "d9d_ops.h"

namespace d9d {
namespace {

template <int VEC>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i += 2) {
    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p + i));
    v[i] = f.x; v[i + 1] = f.y;
  }
}

template <int VEC>
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i += 2) *reinterpret_cast<uint32_t*>(p + i) = pack_bf16x2(v[i], v[i + 1]);
}

// out = R(theta) v for the first rope_dim dims (inverse: R(-theta)); HALF pairs (e, e + rope_dim/2), INTERLEAVED (2i, 2i+1)
template <int VEC>
__device__ __forceinline__ void rotate(const float (&v)[VEC], float (&o)[VEC], const float* __restrict__ cp,
                                       const float* __restrict__ sp, int lane, int rope_dim, int style, bool inverse) {
  if (style == 0) {
    const int half = rope_dim >> 1;
    const int lane_shift = half / VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int e = lane * VEC + i;
      const int src_lane = (e < half) ? lane + lane_shift : lane - lane_shift;
      const float partner = __shfl_sync(0xffffffffu, v[i], src_lane & 31);
      if (e < rope_dim) {
        float s = sp[e];
        if (inverse) s = -s;
        o[i] = v[i] * cp[e] + ((e < half) ? -partner : partner) * s;
      } else {
        o[i] = v[i];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 2) {
      const int e = lane * VEC + i;
      if (e < rope_dim) {
        float s0 = sp[e], s1 = sp[e + 1];
        if (inverse) { s0 = -s0; s1 = -s1; }
        o[i] = v[i] * cp[e] - v[i + 1] * s0;
        o[i + 1] = v[i + 1] * cp[e + 1] + v[i] * s1;
      } else {
        o[i] = v[i]; o[i + 1] = v[i + 1];
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) qk_norm_rope_fwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ wq,
    const __nv_bfloat16* __restrict__ wk, const float* __restrict__ cos_t, const float* __restrict__ sin_t, long long T,
    int Hq, int Hk, int rope_dim, long long ldq, long long ldk, float eps, bool zero_centered, int style,
    __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_out, float* __restrict__ inv_rms) {
  constexpr int D = VEC * 32;
  const int lane = threadIdx.x & 31;
  const int Ht = Hq + Hk;
  const long long total = T * Ht;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  float wqv[VEC], wkv[VEC];
  load_vec<VEC>(wq + lane * VEC, wqv);
  load_vec<VEC>(wk + lane * VEC, wkv);
  if (zero_centered) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { wqv[i] += 1.f; wkv[i] += 1.f; }
  }
  for (long long idx = warp; idx < total; idx += warps) {
    const long long t = idx / Ht;
    const int h = static_cast<int>(idx - t * Ht);
    const bool is_k = h >= Hq;
    const __nv_bfloat16* src = is_k ? k + t * ldk + static_cast<long long>(h - Hq) * D : q + t * ldq + static_cast<long long>(h) * D;
    __nv_bfloat16* dst = is_k ? k_out + (t * Hk + (h - Hq)) * D : q_out + (t * Hq + h) * D;
    float v[VEC], y[VEC], o[VEC];
    load_vec<VEC>(src + lane * VEC, v);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) ss += v[i] * v[i];
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / D + eps);
#pragma unroll
    for (int i = 0; i < VEC; ++i) y[i] = v[i] * r * (is_k ? wkv[i] : wqv[i]);
    rotate<VEC>(y, o, cos_t + t * rope_dim, sin_t + t * rope_dim, lane, rope_dim, style, false);
    store_vec<VEC>(dst + lane * VEC, o);
    if (lane == 0) inv_rms[idx] = r;
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) qk_norm_rope_bwd_kernel(
    const __nv_bfloat16* __restrict__ dq_out, const __nv_bfloat16* __restrict__ dk_out, const __nv_bfloat16* __restrict__ q,
    const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
    const float* __restrict__ cos_t, const float* __restrict__ sin_t, const float* __restrict__ inv_rms, long long T, int Hq,
    int Hk, int rope_dim, long long ldq, long long ldk, bool zero_centered, int style, __nv_bfloat16* __restrict__ dq,
    __nv_bfloat16* __restrict__ dk, float* __restrict__ dw_partial /* [grid, 2, D] */) {
  constexpr int D = VEC * 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int Ht = Hq + Hk;
  const long long total = T * Ht;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  float wqv[VEC], wkv[VEC], accq[VEC], acck[VEC];
  load_vec<VEC>(wq + lane * VEC, wqv);
  load_vec<VEC>(wk + lane * VEC, wkv);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (zero_centered) { wqv[i] += 1.f; wkv[i] += 1.f; }
    accq[i] = 0.f; acck[i] = 0.f;
  }
  for (long long idx = warp; idx < total; idx += warps) {
    const long long t = idx / Ht;
    const int h = static_cast<int>(idx - t * Ht);
    const bool is_k = h >= Hq;
    const __nv_bfloat16* xs = is_k ? k + t * ldk + static_cast<long long>(h - Hq) * D : q + t * ldq + static_cast<long long>(h) * D;
    const long long off = is_k ? (t * Hk + (h - Hq)) * D : (t * Hq + h) * D;
    const __nv_bfloat16* gs = (is_k ? dk_out : dq_out) + off;
    float g[VEC], dy[VEC], x[VEC], dx[VEC];
    load_vec<VEC>(gs + lane * VEC, g);
    load_vec<VEC>(xs + lane * VEC, x);
    rotate<VEC>(g, dy, cos_t + t * rope_dim, sin_t + t * rope_dim, lane, rope_dim, style, true);
    const float r = inv_rms[idx];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float xhat = x[i] * r;
      const float gw = dy[i] * (is_k ? wkv[i] : wqv[i]);
      dot += gw * xhat;
      if (is_k) acck[i] += dy[i] * xhat; else accq[i] += dy[i] * xhat;
      dx[i] = gw;   // finished below once the row mean is known
      x[i] = xhat;
    }
    dot = warp_sum(dot) / D;
#pragma unroll
    for (int i = 0; i < VEC; ++i) dx[i] = r * (dx[i] - x[i] * dot);
    store_vec<VEC>((is_k ? dk : dq) + off + lane * VEC, dx);
  }
  // block-level reduction of the weight gradients: [8 warps][2][D] in shared memory -> one partial row per block
  __shared__ float red[8][2][D];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    red[wid][0][lane * VEC + i] = accq[i];
    red[wid][1][lane * VEC + i] = acck[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * D; e += blockDim.x) {
    const int which = e / D, d = e - which * D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][which][d];
    dw_partial[(static_cast<long long>(blockIdx.x) * 2 + which) * D + d] = s;
  }
}

__global__ void __launch_bounds__(256) qk_dw_reduce_kernel(const float* __restrict__ partial, int blocks, int D,
                                                           float* __restrict__ dwq, float* __restrict__ dwk) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * D) return;
  const int which = e / D, d = e - which * D;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[(static_cast<long long>(b) * 2 + which) * D + d];
  (which ? dwk : dwq)[d] = s;
}

inline int sms() {
  static int n = 0;
  if (!n) { int d; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); }
  return n;
}

inline void check_dims(int D, int rope_dim) {
  const int vec = D / 32;
  if ((D != 64 && D != 128 && D != 256) || rope_dim > D || rope_dim % 2 != 0 || (rope_dim / 2) % vec != 0)
    throw std::runtime_error("d9d qk_norm_rope: head_dim must be 64/128/256 and rope_dim/2 a multiple of head_dim/32");
}

}  // namespace

int qk_norm_rope_grid(long long T, int Ht) {
  long long blocks = (T * Ht + 7) / 8;
  const long long cap = static_cast<long long>(sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks > 0 ? blocks : 1);
}

void qk_norm_rope_fwd(const void* q, const void* k, const void* wq, const void* wk, const float* cos_t, const float* sin_t,
                      long long T, int Hq, int Hk, int D, int rope_dim, long long ldq, long long ldk, float eps,
                      bool zero_centered, int style, void* q_out, void* k_out, float* inv_rms, cudaStream_t s) {
  if (T == 0) return;
  check_dims(D, rope_dim);
  const int grid = qk_norm_rope_grid(T, Hq + Hk);
#define D9D_QKF(V)                                                                                                   \
  qk_norm_rope_fwd_kernel<V><<<grid, 256, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)wq, \
      (const __nv_bfloat16*)wk, cos_t, sin_t, T, Hq, Hk, rope_dim, ldq, ldk, eps, zero_centered, style,                 \
      (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_out, inv_rms)
  if (D == 64) D9D_QKF(2); else if (D == 128) D9D_QKF(4); else D9D_QKF(8);
#undef D9D_QKF
}

void qk_norm_rope_bwd(const void* dq_out, const void* dk_out, const void* q, const void* k, const void* wq, const void* wk,
                      const float* cos_t, const float* sin_t, const float* inv_rms, long long T, int Hq, int Hk, int D,
                      int rope_dim, long long ldq, long long ldk, bool zero_centered, int style, void* dq, void* dk,
                      float* dw_partial, float* dwq, float* dwk, cudaStream_t s) {
  if (T == 0) return;
  check_dims(D, rope_dim);
  const int grid = qk_norm_rope_grid(T, Hq + Hk);
#define D9D_QKB(V)                                                                                                   \
  qk_norm_rope_bwd_kernel<V><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dq_out, (const __nv_bfloat16*)dk_out,        \
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)wq, (const __nv_bfloat16*)wk, cos_t, sin_t, \
      inv_rms, T, Hq, Hk, rope_dim, ldq, ldk, zero_centered, style, (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, dw_partial)
  if (D == 64) D9D_QKB(2); else if (D == 128) D9D_QKB(4); else D9D_QKB(8);
#undef D9D_QKB
  qk_dw_reduce_kernel<<<(2 * D + 255) / 256, 256, 0, s>>>(dw_partial, grid, D, dwq, dwk);
}

}  // namespace d9d
