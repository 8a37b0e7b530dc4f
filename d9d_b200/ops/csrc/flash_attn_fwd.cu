// d9d_b200 — flash-attention forward on tcgen05 (sm_100a).
//
//   O[b, s, h, :] = softmax(scale * Q[b, s, h, :] · K[b, :, h / g, :]^T + mask) · V[b, :, h / g, :]       (GQA, g = Hq / Hk)
//
// One CTA per (128-row query tile, head, batch).  Warp roles (192 threads):
//   warps 0-3  softmax: thread r owns query row r — tcgen05.ld of its S row, online softmax in registers, P (bf16)
//              written to 128B-swizzled shared memory as the A operand of the second GEMM, O rescaled in TMEM
//   warp 4     TMA producer: Q once, then a two-stage ring of K / V tiles (128 keys each)
//   warp 5     MMA issuer:   S = Q·K^T  (UMMA 128x128xD, both operands K-major)          -> TMEM columns [0, 128)
//                            O += P·V   (UMMA 128xDx128, V consumed MN-major, no transpose) -> TMEM columns [128, 128 + D)
// S(j+1) is issued as soon as the softmax warps have pulled S(j) into registers, so the tensor core computes the next
// score tile while the exponentials of the current one are evaluated.
// Layout: q/k/v/o are [B, S, H, D] (the projection outputs, no transposes); LSE is [B, H, S] fp32 (natural log).
#include <stdexcept>
#include <string>

#include "gemm_host.cuh"

namespace d9d {
namespace {

constexpr int FA_BLOCK_M = 128;
constexpr int FA_BLOCK_N = 128;
constexpr int FA_THREADS = 192;
constexpr int UMMA_K_FA = 16;

struct FaParams {
  int B, Sq, Sk, Hq, Hk, causal;
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  float* lse;
};

template <int D>
struct FaCfg {
  static constexpr int Q_BYTES = FA_BLOCK_M * D * 2;
  static constexpr int KV_BYTES = FA_BLOCK_N * D * 2;
  static constexpr int P_BYTES = FA_BLOCK_M * FA_BLOCK_N * 2;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * 2 * KV_BYTES + P_BYTES + 1024 + 256;
  static constexpr uint32_t TMEM_COLS = (128 + D) <= 256 ? 256 : 512;
};

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const FaParams p) {
  using C = FaCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem_q + C::Q_BYTES;                // [2 stages][KV_BYTES]
  uint8_t* smem_v = smem_k + 2 * C::KV_BYTES;           // [2 stages][KV_BYTES]
  uint8_t* smem_p = smem_v + 2 * C::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + C::P_BYTES);
  uint64_t* bar_q = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* o_done = bars + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hk);
  const int q0 = q_tile * FA_BLOCK_M;
  int n_tiles = (p.Sk + FA_BLOCK_N - 1) / FA_BLOCK_N;
  if (p.causal) {
    // bottom-right aligned causal mask: query i sees keys <= i + (Sk - Sq)
    const int last_key = min(p.Sk - 1, q0 + FA_BLOCK_M - 1 + (p.Sk - p.Sq));
    n_tiles = last_key < 0 ? 0 : last_key / FA_BLOCK_N + 1;
  }

  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_ready, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base;         // 128 fp32 columns
  const uint32_t tmem_o = tmem_base + 128;   // D fp32 columns

  if (warp == 4) {
    // ================= TMA producer =================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, C::Q_BYTES);
#pragma unroll
      for (int dc = 0; dc < D / 64; ++dc) tma_load_4d(smem_q + dc * (FA_BLOCK_M * 128), &tmap_q, bar_q, dc * 64, h, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * C::KV_BYTES);
#pragma unroll
        for (int dc = 0; dc < D / 64; ++dc) {
          tma_load_4d(smem_k + st * C::KV_BYTES + dc * (FA_BLOCK_N * 128), &tmap_k, &kv_full[st], dc * 64, hk, j * FA_BLOCK_N, b);
          tma_load_4d(smem_v + st * C::KV_BYTES + dc * (FA_BLOCK_N * 128), &tmap_v, &kv_full[st], dc * 64, hk, j * FA_BLOCK_N, b);
        }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = make_idesc_bf16(FA_BLOCK_M, FA_BLOCK_N, false, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(FA_BLOCK_M, D, false, true);
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      if (j > 0) mbar_wait(s_free, (j - 1) & 1);  // softmax pulled S(j-1) into registers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t qa = smem_u32(smem_q), ka = smem_u32(smem_k + st * C::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < D / UMMA_K_FA; ++kk) {
          const uint32_t off = (kk / 4) * (FA_BLOCK_M * 128) + (kk % 4) * 32;  // 64-element boxes, 32 B per UMMA_K step
          umma_f16(tmem_s, make_smem_desc_sw128(qa + off, 16, 1024), make_smem_desc_sw128(ka + off, 16, 1024), idesc_s, kk != 0);
        }
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_tiles > 0) {
      mbar_wait(bar_q, 0);
      issue_s(0);
    }
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_s(j + 1);
      const int st = j & 1;
      mbar_wait(p_ready, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pa = smem_u32(smem_p), va = smem_u32(smem_v + st * C::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < FA_BLOCK_N / UMMA_K_FA; ++kk) {
          const uint64_t da = make_smem_desc_sw128(pa + (kk / 4) * (FA_BLOCK_M * 128) + (kk % 4) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(va + kk * (UMMA_K_FA * 128), FA_BLOCK_N * 128, 1024);
          umma_f16(tmem_o, da, db, idesc_o, (j | kk) != 0);
        }
        umma_commit(o_done);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ================= softmax / correction / epilogue (thread = query row) =================
    const int row = warp * 32 + lane;
    const int qi = q0 + row;
    const int key_limit = p.causal ? qi + (p.Sk - p.Sq) : p.Sk - 1;  // largest visible key index
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[FA_BLOCK_N];
#pragma unroll
      for (int c = 0; c < FA_BLOCK_N / 32; ++c) {
        uint32_t (&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]);
        tmem_ld_32x32b_x32(tmem_s + lane_base + c * 32, chunk);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      // mask + running max (in the log2 domain: x = s * scale * log2(e))
      const int k0 = j * FA_BLOCK_N;
      float mx = m;
#pragma unroll
      for (int i = 0; i < FA_BLOCK_N; ++i) {
        float x = __uint_as_float(s[i]) * p.scale_log2;
        if (k0 + i > key_limit || k0 + i >= p.Sk) x = -INFINITY;
        s[i] = __float_as_uint(x);
        mx = fmaxf(mx, x);
      }
      const float m_safe = (mx == -INFINITY) ? 0.f : mx;
      const float alpha = exp2f(m - m_safe);  // m = -inf on the first tile -> 0
      float sum = 0.f;
      uint32_t packed[FA_BLOCK_N / 2];
#pragma unroll
      for (int i = 0; i < FA_BLOCK_N; i += 2) {
        const float p0 = exp2f(__uint_as_float(s[i]) - m_safe), p1 = exp2f(__uint_as_float(s[i + 1]) - m_safe);
        sum += p0 + p1;
        packed[i >> 1] = pack_bf16x2(p0, p1);
      }
      l = l * alpha + sum;
      m = mx;
      if (j > 0) mbar_wait(o_done, (j - 1) & 1);  // P buffer is free and O holds tiles < j
      tc_fence_after();
      // P row -> shared memory (K-major, 128B swizzle: two 64-key boxes of 128 rows x 128 B)
#pragma unroll
      for (int kb = 0; kb < FA_BLOCK_N / 64; ++kb) {
        uint8_t* my_row = smem_p + kb * (FA_BLOCK_M * 128) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(my_row + ((c ^ (row & 7)) << 4)) =
              make_uint4(packed[kb * 32 + 4 * c], packed[kb * 32 + 4 * c + 1], packed[kb * 32 + 4 * c + 2], packed[kb * 32 + 4 * c + 3]);
      }
      // rescale the running output when the row maximum moved (skipped warp-wide when nobody's did)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tmem_o + lane_base + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(tmem_o + lane_base + c * 32, o);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> global, LSE
    const bool row_ok = qi < p.Sq;
    if (n_tiles > 0) {
      mbar_wait(o_done, (n_tiles - 1) & 1);
      tc_fence_after();
    }
    const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
    __nv_bfloat16* dst = p.out + ((static_cast<long long>(b) * p.Sq + qi) * p.Hq + h) * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      if (n_tiles > 0) {
        tmem_ld_32x32b_x32(tmem_o + lane_base + c * 32, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0;
      }
      if (row_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[v * 8 + 0]) * inv_l, __uint_as_float(o[v * 8 + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[v * 8 + 2]) * inv_l, __uint_as_float(o[v * 8 + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[v * 8 + 4]) * inv_l, __uint_as_float(o[v * 8 + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[v * 8 + 6]) * inv_l, __uint_as_float(o[v * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = u;
        }
      }
    }
    if (row_ok && p.lse != nullptr)
      p.lse[(static_cast<long long>(b) * p.Hq + h) * p.Sq + qi] = (l > 0.f) ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
    tc_fence_before();
  }

  __syncthreads();
  tc_fence_after();
  if (warp == 5) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// [B, S, H, D] bf16 -> 4-D tensor map (d, h, s, b), box = 64 x 1 x 128 x 1, 128B swizzle
CUtensorMap make_tmap_bshd(const void* base, int B, int S, int H, int D) {
  CUtensorMap m;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(D) * 2, static_cast<cuuint64_t>(H) * D * 2, static_cast<cuuint64_t>(S) * H * D * 2};
  cuuint32_t box[4] = {64, 1, 128, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) throw std::runtime_error("d9d flash_attn: tensors must be 16-byte aligned");
  CUresult r = gemm::get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("d9d flash_attn: cuTensorMapEncodeTiled failed: " + std::to_string(r));
  return m;
}

template <int D>
void launch_fa(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Sq, int Sk, int Hq, int Hk,
               float scale, bool causal, cudaStream_t stream) {
  using C = FaCfg<D>;
  auto kern = flash_attn_fwd_kernel<D>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    configured = true;
  }
  const CUtensorMap tq = make_tmap_bshd(q, B, Sq, Hq, D), tk = make_tmap_bshd(k, B, Sk, Hk, D), tv = make_tmap_bshd(v, B, Sk, Hk, D);
  FaParams p{B, Sq, Sk, Hq, Hk, causal ? 1 : 0, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out), lse};
  dim3 grid((Sq + FA_BLOCK_M - 1) / FA_BLOCK_M, Hq, B);
  kern<<<grid, FA_THREADS, C::SMEM_BYTES, stream>>>(tq, tk, tv, p);
}

}  // namespace

void flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Sq, int Sk, int Hq, int Hk,
                    int D, float scale, bool causal, cudaStream_t stream) {
  if (B == 0 || Sq == 0 || Hq == 0) return;
  if (Hq % Hk != 0) throw std::runtime_error("d9d flash_attn: Hq must be a multiple of Hk");
  if (D == 64) launch_fa<64>(q, k, v, out, lse, B, Sq, Sk, Hq, Hk, scale, causal, stream);
  else if (D == 128) launch_fa<128>(q, k, v, out, lse, B, Sq, Sk, Hq, Hk, scale, causal, stream);
  else throw std::runtime_error("d9d flash_attn: head_dim must be 64 or 128");
}

}  // namespace d9d
