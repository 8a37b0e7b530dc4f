// d9d_b200 — persistent warp-specialised tcgen05 GEMM for 8-bit floating point operands (sm_100a).
//
//   D[M,N] (bf16) = scale( A[M,K] · B[N,K]^T )       A, B: e4m3, K-major; fp32 accumulation in TMEM.
//
// Two scaling recipes share the mainloop:
//   ROWCOL  kind::f8f6f4        D = acc * scale_a[row] * scale_b[col]  (per-token / per-output-channel or per-tensor scales,
//                               applied in the epilogue while the accumulator is read back from tensor memory)
//   MX      kind::mxf8f6f4.block_scale   one UE8M0 scale per 32 K-elements of every row of A and B (OCP MXFP8).  The scale
//                               factors travel with the operand stages: a 512-byte block (128 rows x 4 K-groups) per operand
//                               tile is bulk-copied into shared memory next to the A/B tiles and moved into tensor memory by
//                               tcgen05.cp right before the four K=32 MMAs that consume it.
// Same structure as the bf16 kernel (gemm_sm100.cuh): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue
// (TMEM -> registers -> swizzled smem -> TMA store); a stage holds 128 bytes of K (= 128 fp8 elements = 4 MMAs).
#pragma once

#include "common.cuh"

namespace d9d {
namespace fp8 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 128;  // fp8 elements = bytes = one 128B swizzle row
constexpr int UMMA_K = 32;
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_WARPS = 4;
constexpr int SF_BLOCK_BYTES = 512;  // 128 rows x 4 K-groups of 32 elements

enum Scaling : int { ROWCOL = 0, MX = 1 };

struct Params {
  int M, N, K;
  const float* scale_a;  // ROWCOL: [M] or nullptr (=1)
  const float* scale_b;  // ROWCOL: [N] or nullptr (=1)
  float scale_scalar;    // ROWCOL: extra scalar factor
  const uint8_t* sfa;    // MX: [ceil(M/128), K/128, 512] UE8M0 blocks, byte (r%32)*16 + (r/32)*4 + kgroup
  const uint8_t* sfb;    // MX: [ceil(N/128), K/128, 512]
};

template <int BLOCK_N, int SCALING>
struct Cfg {
  static constexpr bool IS_MX = (SCALING == MX);
  static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K;
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K;
  static constexpr int SFA_STAGE_BYTES = IS_MX ? SF_BLOCK_BYTES : 0;
  static constexpr int SFB_BLOCKS = (BLOCK_N + 127) / 128;
  static constexpr int SFB_STAGE_BYTES = IS_MX ? SFB_BLOCKS * SF_BLOCK_BYTES : 0;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES + SFA_STAGE_BYTES + SFB_STAGE_BYTES;
  // tensor memory: fp32 accumulators + (MX) 4 columns per 128-row scale block.  A 256-wide tile leaves no room for a second
  // accumulator next to the scale factors, so the block-scaled kernel runs single-buffered accumulators.
  static constexpr int ACC_STAGES = IS_MX ? (BLOCK_N <= 128 ? 2 : 1) : 2;
  static constexpr int SF_COLS = IS_MX ? 4 * (1 + SFB_BLOCKS) : 0;
  static constexpr int TMEM_COLS_NEEDED = ACC_STAGES * BLOCK_N + SF_COLS;
  static_assert(TMEM_COLS_NEEDED <= 512, "tile does not fit tensor memory");
  static constexpr uint32_t TMEM_COLS = TMEM_COLS_NEEDED <= 32    ? 32
                                        : TMEM_COLS_NEEDED <= 64  ? 64
                                        : TMEM_COLS_NEEDED <= 128 ? 128
                                        : TMEM_COLS_NEEDED <= 256 ? 256
                                                                  : 512;
  static constexpr int SFA_COL = ACC_STAGES * BLOCK_N;
  static constexpr int SFB_COL = SFA_COL + 4;
  static constexpr int EPI_BUFS = IS_MX ? 1 : 2;
  static constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * EPI_BUFS * 4096;
  static constexpr int SMEM_BUDGET = 227 * 1024 - 2048 - 256 - EPI_STAGING_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static_assert(STAGES >= 3, "need at least 3 operand stages");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 2048 /*two alignments*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tile_coord(int tile, int m_tiles, int n_tiles, int& m_blk, int& n_blk) {
  constexpr int GROUP_M = 8;  // L2-friendly rasterisation (same as the bf16 kernel)
  const int group_span = GROUP_M * n_tiles;
  const int gid = tile / group_span;
  const int first_m = gid * GROUP_M;
  const int gm = min(m_tiles - first_m, GROUP_M);
  const int r = tile - gid * group_span;
  m_blk = first_m + (r % gm);
  n_blk = r / gm;
}

template <int BLOCK_N, int SCALING>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_d, const Params p) {
  using C = Cfg<BLOCK_N, SCALING>;
  constexpr int STAGES = C::STAGES;
  constexpr int ACC_STAGES = C::ACC_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + STAGES * C::A_STAGE_BYTES;
  uint8_t* smem_sfa = smem_b + STAGES * C::B_STAGE_BYTES;
  uint8_t* smem_sfb = smem_sfa + STAGES * C::SFA_STAGE_BYTES;
  // the swizzled epilogue staging boxes must start on a 1024-byte boundary (the scale-factor regions are 512-byte granular)
  uint8_t* smem_epi = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_sfb + STAGES * C::SFB_STAGE_BYTES) + 1023) & ~uintptr_t(1023));
  static_assert((C::A_STAGE_BYTES % 1024) == 0 && (C::B_STAGE_BYTES % 1024) == 0, "operand stages keep the 1024-byte swizzle alignment");
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int sf_k_blocks = k_blocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coord(tile, m_tiles, n_tiles, m_blk, n_blk);
        const int m_idx = m_blk * BLOCK_M, n_idx = n_blk * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_2d(smem_a + stage * C::A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_idx);
          tma_load_2d(smem_b + stage * C::B_STAGE_BYTES, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_idx);
          if constexpr (C::IS_MX) {
            bulk_load_1d(smem_sfa + stage * C::SFA_STAGE_BYTES,
                         p.sfa + (static_cast<long long>(m_blk) * sf_k_blocks + kb) * SF_BLOCK_BYTES, SF_BLOCK_BYTES, &full_bar[stage]);
#pragma unroll
            for (int j = 0; j < C::SFB_BLOCKS; ++j)
              bulk_load_1d(smem_sfb + stage * C::SFB_STAGE_BYTES + j * SF_BLOCK_BYTES,
                           p.sfb + (static_cast<long long>(n_blk * C::SFB_BLOCKS + j) * sf_k_blocks + kb) * SF_BLOCK_BYTES,
                           SF_BLOCK_BYTES, &full_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_STAGE_BYTES);
          if constexpr (C::IS_MX) {
            // scale factors of this stage: smem -> tmem (same pipe as the MMAs below, which therefore see them)
            tmem_cp_32x128b_warpx4(tmem_base + C::SFA_COL,
                                   make_smem_desc_nosw(smem_u32(smem_sfa + stage * C::SFA_STAGE_BYTES), 16, 128));
#pragma unroll
            for (int j = 0; j < C::SFB_BLOCKS; ++j)
              tmem_cp_32x128b_warpx4(tmem_base + C::SFB_COL + 4 * j,
                                     make_smem_desc_nosw(smem_u32(smem_sfb + stage * C::SFB_STAGE_BYTES + j * SF_BLOCK_BYTES), 16, 128));
          }
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K, 16, 1024);
            if constexpr (C::IS_MX) {
              // K-group k of the stage = byte k of every 32-bit scale word: selected through the sf ids of the instruction
              // descriptor (mirrored in the two top bits of the tensor-memory address)
              const uint32_t sel = static_cast<uint32_t>(k) << 30;
              umma_mxf8(tmem_d, da, db, make_idesc_mxe4m3(BLOCK_M, BLOCK_N, k, k), (kb | k) != 0,
                        (tmem_base + C::SFA_COL) | sel, (tmem_base + C::SFB_COL) | sel);
            } else {
              umma_f8(tmem_d, da, db, make_idesc_e4m3(BLOCK_M, BLOCK_N), (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue (4 warps, one TMEM lane quadrant each) =================
    const int quad = warp & 3;
    const int lane = lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t epi_chunk = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coord(tile, m_tiles, n_tiles, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + quad * 32 + lane;
      const int col0 = n_blk * BLOCK_N;
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
      float row_scale = 1.f;
      if constexpr (!C::IS_MX) {
        row_scale = p.scale_scalar;
        if (p.scale_a != nullptr && row < p.M) row_scale *= p.scale_a[row];
      }
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 64; ++c) {
        const int cbase = col0 + c * 64;
        if (cbase >= p.N) break;  // warp-uniform
        uint8_t* stage_buf = smem_epi + (quad * C::EPI_BUFS + (epi_chunk % C::EPI_BUFS)) * 4096;
        ++epi_chunk;
        if (lane == 0) tma_store_wait_read<C::EPI_BUFS - 1>();
        __syncwarp();
        uint32_t packed[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 64 + h * 32, r);
          tmem_ld_wait();
          if constexpr (!C::IS_MX) {
            // lane l holds the scale of column cbase + 32h + l; broadcast while scaling
            float my_col_scale = 1.f;
            if (p.scale_b != nullptr && cbase + h * 32 + lane < p.N) my_col_scale = p.scale_b[cbase + h * 32 + lane];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float cs = __shfl_sync(0xffffffffu, my_col_scale, i);
              r[i] = __float_as_uint(__uint_as_float(r[i]) * row_scale * cs);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i)
            packed[h * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        }
        uint8_t* my_row = stage_buf + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(my_row + ((j ^ (lane & 7)) << 4)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmap_d, stage_buf, cbase, m_blk * BLOCK_M + quad * 32, 0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

}  // namespace fp8
}  // namespace d9d
