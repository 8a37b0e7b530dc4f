// d9d_b200 — data-parallel optimizer step over NVLink / NVSwitch peer memory (no NCCL on the data path).
//
// Every replica keeps its gradients in a *symmetric* fp32 arena and its parameters in a symmetric bf16 arena
// (same layout on every rank, mapped into every peer's address space, plus an NVSwitch multicast mapping).
// Rank r owns the contiguous element range [begin, end) of both arenas and the optimizer state for that range only:
//
//   pass 1  nvl_reduce_shard   own_grad[i] = sum_over_replicas grad[i]      one multimem.ld_reduce per 16 bytes: the
//                                                                           NVSwitch adds the replicas' values in the
//                                                                           fabric, only the reduced vector crosses
//                                                                           this GPU's NVLink; also accumulates
//                                                                           sum(g^2) of the shard for clipping.
//   pass 2  nvl_adamw_shard    AdamW with stochastic rounding on the shard, the updated bf16 parameters are pushed
//                              to *all* replicas with one multimem.st per 16 bytes (the switch replicates).
//
// i.e. reduce-scatter + the optimizer + all-gather collapse into two memory-bound kernels whose NVLink traffic per
// rank is N/W*4 B in and N/W*2 B out (N parameters, W replicas) instead of the 2*(W-1)/W*N*4 B of a ring all-reduce.
// Without a multicast mapping the kernels fall back to explicit peer loads / stores over the peer pointer table.
#include <stdexcept>

#include "common.cuh"
#include "d9d_ops.h"

namespace d9d {
namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float* mc_ptr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_ptr)
               : "memory");
  return v;
}

__device__ __forceinline__ void multimem_st_b32x4(void* mc_ptr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(mc_ptr),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}

__device__ __forceinline__ float4 ld_peer_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__device__ __forceinline__ void st_peer_b32x4(void* p, uint4 v) {
  asm volatile("st.global.relaxed.sys.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// own_grad[begin:end) = sum_r grad_r[begin:end); sumsq += sum of squares of the reduced values.
// UNROLL independent 16-byte requests per thread are issued back to back: NVLink / NVSwitch round trips are several
// microseconds, so throughput is set by the bytes in flight.
template <int UNROLL>
__global__ void __launch_bounds__(512) nvl_reduce_shard_kernel(const float* const* __restrict__ peer_grads,
                                                               const float* __restrict__ mc_grad,
                                                               float* __restrict__ own_grad, long long begin,
                                                               long long end, int world, int rank,
                                                               float* __restrict__ sumsq) {
  const long long n4 = (end - begin) >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * UNROLL) {
    float4 v[UNROLL];
    if (mc_grad != nullptr) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4) v[u] = multimem_ld_reduce_add_f32x4(mc_grad + begin + (i << 2));
      }
    } else {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4) v[u] = *reinterpret_cast<const float4*>(own_grad + begin + (i << 2));
      }
      for (int r = 1; r < world; ++r) {  // start at the next rank so that the peers are not all hit at once
        const float* peer = peer_grads[(rank + r) % world] + begin;
        float4 o[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long long i = i0 + u * stride;
          if (i < n4) o[u] = ld_peer_f32x4(peer + (i << 2));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long long i = i0 + u * stride;
          if (i < n4) { v[u].x += o[u].x; v[u].y += o[u].y; v[u].z += o[u].z; v[u].w += o[u].w; }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        *reinterpret_cast<float4*>(own_grad + begin + (i << 2)) = v[u];
        acc += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
      }
    }
  }
  acc = warp_sum(acc);
  __shared__ float part[16];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(sumsq, v);
  }
}

// AdamW + stochastic rounding on [begin, end); states are shard-local (index 0 == element `begin`).
template <bool STATE_BF16>
__global__ void __launch_bounds__(256) nvl_adamw_shard_kernel(uint16_t* const* __restrict__ peer_params,
                                                              uint16_t* __restrict__ mc_param,
                                                              uint16_t* __restrict__ own_param,
                                                              const float* __restrict__ own_grad,
                                                              void* __restrict__ exp_avg, void* __restrict__ exp_avg_sq,
                                                              long long begin, long long end, int world, int rank,
                                                              float lr, float beta1, float beta2, float eps, float wd,
                                                              float bc1, float bc2, uint64_t seed,
                                                              const float* __restrict__ grad_scale) {
  const float gs = grad_scale ? *grad_scale : 1.f;
  const long long n8 = (end - begin) >> 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long local = i << 3;
    const long long e = begin + local;
    float pf[8], gf[8], mf[8], vf[8];
    {
      const uint4 pv = *reinterpret_cast<const uint4*>(own_param + e);
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
      const float4 g0 = *reinterpret_cast<const float4*>(own_grad + e);
      const float4 g1 = *reinterpret_cast<const float4*>(own_grad + e + 4);
      gf[0] = g0.x; gf[1] = g0.y; gf[2] = g0.z; gf[3] = g0.w; gf[4] = g1.x; gf[5] = g1.y; gf[6] = g1.z; gf[7] = g1.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(pw[j]);
        pf[2 * j] = f.x; pf[2 * j + 1] = f.y;
      }
      if (STATE_BF16) {
        const uint4 mv = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(exp_avg) + local);
        const uint4 vv = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(exp_avg_sq) + local);
        const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w}, vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = unpack_bf16x2(mw[j]), b = unpack_bf16x2(vw[j]);
          mf[2 * j] = a.x; mf[2 * j + 1] = a.y; vf[2 * j] = b.x; vf[2 * j + 1] = b.y;
        }
      } else {
        const float* m = static_cast<const float*>(exp_avg) + local;
        const float* v = static_cast<const float*>(exp_avg_sq) + local;
#pragma unroll
        for (int j = 0; j < 8; ++j) { mf[j] = m[j]; vf[j] = v[j]; }
      }
    }
    const uint64_t ctr = static_cast<uint64_t>(e >> 3);  // counter = global arena position: independent of sharding
    const uint4 rp = philox4x32(seed, ctr);
    uint4 rm = make_uint4(0, 0, 0, 0), rv = make_uint4(0, 0, 0, 0);
    if (STATE_BF16) {
      rm = philox4x32(seed + 42, ctr);
      rv = philox4x32(seed + 67, ctr);
    }
    const uint32_t rpa[4] = {rp.x, rp.y, rp.z, rp.w}, rma[4] = {rm.x, rm.y, rm.z, rm.w}, rva[4] = {rv.x, rv.y, rv.z, rv.w};
    uint16_t po[8], mo[8], vo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = gf[j] * gs;
      float pw = pf[j] * (1.f - lr * wd);
      const float mn = beta1 * mf[j] + (1.f - beta1) * g;
      const float vn = beta2 * vf[j] + (1.f - beta2) * (g * g);
      pw -= (lr * (mn / bc1)) / (sqrtf(vn / bc2) + eps);
      const uint32_t sel = (j & 1) ? 16 : 0;
      po[j] = sr_bf16_bits(pw, (rpa[j >> 1] >> sel) & 0xFFFF);
      mf[j] = mn; vf[j] = vn;
      if (STATE_BF16) {
        mo[j] = sr_bf16_bits(mn, (rma[j >> 1] >> sel) & 0xFFFF);
        vo[j] = sr_bf16_bits(vn, (rva[j >> 1] >> sel) & 0xFFFF);
      }
    }
    const uint4 pnew = *reinterpret_cast<const uint4*>(po);
    if (mc_param != nullptr) {
      multimem_st_b32x4(mc_param + e, pnew);  // one store, replicated to every replica by the switch
    } else {
      *reinterpret_cast<uint4*>(own_param + e) = pnew;
      for (int r = 1; r < world; ++r) st_peer_b32x4(peer_params[(rank + r) % world] + e, pnew);
    }
    if (STATE_BF16) {
      *reinterpret_cast<uint4*>(static_cast<uint16_t*>(exp_avg) + local) = *reinterpret_cast<const uint4*>(mo);
      *reinterpret_cast<uint4*>(static_cast<uint16_t*>(exp_avg_sq) + local) = *reinterpret_cast<const uint4*>(vo);
    } else {
      float* m = static_cast<float*>(exp_avg) + local;
      float* v = static_cast<float*>(exp_avg_sq) + local;
#pragma unroll
      for (int j = 0; j < 8; ++j) { m[j] = mf[j]; v[j] = vf[j]; }
    }
  }
}

inline int sms() {
  static int n = 0;
  if (!n) { int d; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); }
  return n;
}

}  // namespace

void nvl_reduce_shard(const float* const* peer_grads, const float* mc_grad, float* own_grad, long long begin,
                      long long end, int world, int rank, float* sumsq, cudaStream_t stream) {
  if (end <= begin) return;
  if (((begin | end) & 7) != 0) throw std::runtime_error("d9d nvl_reduce_shard: shard bounds must be multiples of 8 elements");
  const long long n4 = (end - begin) >> 2;
  long long blocks = (n4 + 512 * 4 - 1) / (512 * 4);
  const long long cap = static_cast<long long>(sms()) * 4;
  if (blocks > cap) blocks = cap;
  nvl_reduce_shard_kernel<4><<<static_cast<int>(blocks), 512, 0, stream>>>(peer_grads, mc_grad, own_grad, begin, end,
                                                                           world, rank, sumsq);
}

void nvl_adamw_shard(void* const* peer_params, void* mc_param, void* own_param, const float* own_grad, void* exp_avg,
                     void* exp_avg_sq, long long begin, long long end, int world, int rank, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, uint64_t seed,
                     const float* grad_scale, bool state_bf16, cudaStream_t stream) {
  if (end <= begin) return;
  if (((begin | end) & 7) != 0) throw std::runtime_error("d9d nvl_adamw_shard: shard bounds must be multiples of 8 elements");
  const long long n8 = (end - begin) >> 3;
  long long blocks = (n8 + 255) / 256;
  const long long cap = static_cast<long long>(sms()) * 8;
  if (blocks > cap) blocks = cap;
  auto pp = reinterpret_cast<uint16_t* const*>(peer_params);
  if (state_bf16)
    nvl_adamw_shard_kernel<true><<<static_cast<int>(blocks), 256, 0, stream>>>(
        pp, static_cast<uint16_t*>(mc_param), static_cast<uint16_t*>(own_param), own_grad, exp_avg, exp_avg_sq, begin,
        end, world, rank, lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, seed, grad_scale);
  else
    nvl_adamw_shard_kernel<false><<<static_cast<int>(blocks), 256, 0, stream>>>(
        pp, static_cast<uint16_t*>(mc_param), static_cast<uint16_t*>(own_param), own_grad, exp_avg, exp_avg_sq, begin,
        end, world, rank, lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, seed, grad_scale);
}

}  // namespace d9d
