// d9d_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = A[M,K] · B[N,K]^T      bf16 operands, fp32 accumulation in TMEM.
//
// * operands are staged by TMA (128B swizzle) into a multi-stage smem ring,
// * one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16),
// * accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1,
// * either operand may be K-major or MN-major (needed for dgrad / wgrad without transposes),
// * three problem shapes share the mainloop:
//     DENSE      plain GEMM
//     GROUPED_M  rows of A/D are grouped by expert (128-row aligned segments, device-side tile->expert table),
//                B is [E, ...]  -> MoE forward / dgrad without any host sync
//     GROUPED_K  the reduction dim is grouped (per-expert weight gradients), D is [E, M, N]
// * epilogues: bf16 store, fp32 store, fp32/bf16 accumulate, and the fused linear-cross-entropy epilogues.
//   Output tiles leave the SM through TMA: TMEM -> registers -> 128B-swizzled smem staging -> cp.async.bulk.tensor
//   store, or cp.reduce.async.bulk.tensor (.add) for the accumulate epilogues, so `D += acc` costs no global loads
//   in the SM (the add happens in L2) and partial tiles are clipped by the tensor map.
// * split-K (DENSE, accumulate epilogues): several CTAs reduce-add K-slices of one output tile; used when a wgrad
//   has few output tiles but a long reduction dim.
#pragma once

#include "common.cuh"

namespace d9d {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;  // warp0: TMA, warp1: MMA, warps2-5: epilogue
constexpr int NUM_EPI_WARPS = 4;

enum Mode : int { DENSE = 0, GROUPED_M = 1, GROUPED_K = 2 };
enum Epi : int {
  EPI_BF16 = 0,       // D(bf16) = acc
  EPI_F32 = 1,        // D(f32) = acc
  EPI_F32_ACC = 2,    // D(f32) += acc
  EPI_BF16_ACC = 3,   // D(bf16) += acc   (fp32 add, rn)
  EPI_CE_LSE = 4,     // fused linear-CE forward: per (row, n_tile) running max / sum-exp / target logit
  EPI_CE_DLOGITS = 5  // fused linear-CE backward: D(bf16) = g[row] * (exp(acc - lse[row]) - [col == target[row]])
};

// Tensor-parallel communication fused into the GEMM (operands / outputs living in NVLink peer memory):
//   COMM_AG_A   A's rows (the M dim) are sharded over the peers: every m-tile is TMA-loaded straight from the owner's
//               shard                                   => all-gather -> GEMM without a gathered copy of A
//   COMM_RS_D   every output tile is TMA reduce-added into the shard of the peer that owns those rows
//                                                        => GEMM -> reduce-scatter without a partial-sum buffer
//   COMM_AG_KA / COMM_AG_KB   A / B (MN-major) is sharded along the *reduction* dim (wgrad of a sequence-parallel
//               layer): every k-block is loaded from the owner's shard
// Row index -> (owner, local row):  blk = idx / comm_block_rows; owner = blk % comm_world;
//                                   local = (blk / comm_world) * comm_block_rows + idx % comm_block_rows
// (covers both "tokens sharded as one contiguous block per rank" and "[batch, seq/W] per rank, batch-major").
//   COMM_WAIT_A A is a *local* gathered buffer that is being filled shard by shard (peer copies running concurrently on
//               the copy engines / another stream); the producer waits on the owner's arrival flag before the first
//               load of an m-tile and tiles are visited starting with the local shard  => all-gather overlapped with
//               the GEMM tile by tile while every remote byte crosses NVLink exactly once
enum Comm : int { COMM_NONE = 0, COMM_AG_A = 1, COMM_RS_D = 2, COMM_AG_KA = 3, COMM_AG_KB = 4, COMM_WAIT_A = 5 };
constexpr int MAX_PEERS = 8;
struct PeerMaps { CUtensorMap m[MAX_PEERS]; };
struct NoPeerMaps {};
template <int COMM> struct PeerArg { using type = PeerMaps; };
template <> struct PeerArg<COMM_NONE> { using type = NoPeerMaps; };
template <> struct PeerArg<COMM_WAIT_A> { using type = NoPeerMaps; };

struct Params {
  int M, N, K;            // DENSE: problem dims. GROUPED_M: M = padded row capacity. GROUPED_K: M,N = per-expert out dims
  int num_groups;         // experts (1 for dense)
  int k_splits;           // DENSE + accumulate epilogue: K is cut into this many slices (1 => off)
  int tma_epilogue;       // 1: output through the tmap_d TMA store/reduce path, 0: direct global stores
  int comm_world;         // COMM_*: number of peers
  int comm_block_rows;    // COMM_*: rows per (rank, block) of the sharded dim
  int comm_rank;          // COMM_WAIT_A: this rank (its shard needs no wait)
  int comm_m_rot;         // COMM_WAIT_A: m-blocks are visited starting at this block (the local shard first)
  const int* comm_flags;  // COMM_WAIT_A: [comm_world] arrival flags of the shards (non-zero = landed)
  void* D;                // output
  long long ldd;          // leading dim of D (elements)
  long long d_group_stride;  // GROUPED_K: elements between per-expert outputs
  const int* tile_group;  // GROUPED_M: [M/128] expert id per m-tile (-1 => unused tile)
  const int* group_offsets;  // GROUPED_K: [E+1] row offsets (multiples of BLOCK_K) into the grouped K dimension
  // fused CE
  const long long* ce_target;  // [M] int64 targets (ignore_index allowed)
  const float* ce_lse;         // [M]           (DLOGITS)
  const float* ce_grad;        // [M]           (DLOGITS) upstream grad per row
  float* ce_part_max;          // [n_tiles, M]  (LSE)
  float* ce_part_sum;          // [n_tiles, M]  (LSE)
  float* ce_tgt_logit;         // [M]           (LSE) written by the tile that owns the target column
  long long ce_ignore_index;
  float ce_softcap;            // logits -> softcap * tanh(logits / softcap) (0 = off)
  const float* ce_bias;        // [N] fp32 added to the logits (or nullptr)
  long long ce_col_offset;     // B is a row slice of the classifier: global class id of its first row
};

// logit transform shared by the fused-CE epilogues: z = softcap(acc + bias); returns z and dz/dacc
__device__ __forceinline__ float ce_transform(float acc, float bias, float cap, float inv_cap, float& dz) {
  float z = acc + bias;
  dz = 1.f;
  if (cap > 0.f) {
    const float th = tanhf(z * inv_cap);
    z = cap * th;
    dz = 1.f - th * th;
  }
  return z;
}

// CTAS = 2: a CTA pair works on a 256 x BLOCK_N tile (cta_group::2): every CTA stages its own 128 rows of A and its own
// BLOCK_N / 2 rows of B, i.e. a third less operand traffic per flop at BLOCK_N = 256 and two more stages in flight.
template <int BLOCK_N, int CTAS = 1>
struct Cfg {
  static_assert(CTAS == 1 || CTAS == 2, "one CTA or a CTA pair");
  static_assert(CTAS == 1 || BLOCK_N % 128 == 0, "a CTA pair splits B into halves of whole 64-column swizzle atoms");
  static constexpr int TILE_M = BLOCK_M * CTAS;
  static constexpr int B_ROWS = BLOCK_N / CTAS;  // rows of B staged by one CTA
  static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // Operand bytes in flight decide the throughput of the mid-size GEMMs (per-SM bandwidth = bytes in flight / load latency),
  // so the 192-wide tile trades the second epilogue staging buffer for a fifth operand stage (200 KiB in flight).
  static constexpr int EPI_BUFS = (BLOCK_N == 192 && CTAS == 1) ? 1 : 2;
  static constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * EPI_BUFS * 4096;  // per epilogue warp: EPI_BUFS x (32 rows x 128 B)
  static constexpr int SMEM_BUDGET = 227 * 1024 - 1024 - 256 - EPI_STAGING_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS_NEEDED = 2 * BLOCK_N;
  static constexpr uint32_t TMEM_COLS = TMEM_COLS_NEEDED <= 32    ? 32
                                        : TMEM_COLS_NEEDED <= 64  ? 64
                                        : TMEM_COLS_NEEDED <= 128 ? 128
                                        : TMEM_COLS_NEEDED <= 256 ? 256
                                                                  : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct TileCoord {
  int m_blk, n_blk, group;
  int k_begin, k_blocks;  // k offset (elements) and number of BLOCK_K chunks
  bool valid;
};

template <int MODE, int BLOCK_N, int TILE_M = BLOCK_M>
__device__ __forceinline__ int num_tiles(const Params& p) {
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  if (MODE == GROUPED_K) return p.num_groups * m_tiles * n_tiles;
  if (MODE == DENSE) return m_tiles * n_tiles * p.k_splits;
  return m_tiles * n_tiles;
}

template <int MODE, int BLOCK_N, int TILE_M = BLOCK_M>
__device__ __forceinline__ TileCoord get_tile(const Params& p, int tile) {
  TileCoord t;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  t.valid = true;
  if (MODE == GROUPED_K) {
    const int per_group = m_tiles * n_tiles;
    t.group = tile / per_group;
    const int r = tile - t.group * per_group;
    t.m_blk = r / n_tiles;
    t.n_blk = r - t.m_blk * n_tiles;
    const int k0 = p.group_offsets[t.group], k1 = p.group_offsets[t.group + 1];
    t.k_begin = k0;
    t.k_blocks = (k1 - k0 + BLOCK_K - 1) / BLOCK_K;
    return t;
  }
  int k_split = 0;
  if (MODE == DENSE && p.k_splits > 1) {  // tile = split * (m_tiles * n_tiles) + mn: one K-slice per wave
    k_split = tile / (m_tiles * n_tiles);
    tile -= k_split * (m_tiles * n_tiles);
  }
  // L2-friendly rasterisation: walk GROUP_M m-blocks for each n-block before moving on.
  constexpr int GROUP_M = 8;
  const int group_span = GROUP_M * n_tiles;
  const int gid = tile / group_span;
  const int first_m = gid * GROUP_M;
  const int gm = min(m_tiles - first_m, GROUP_M);
  const int r = tile - gid * group_span;
  t.m_blk = first_m + (r % gm);
  t.n_blk = r / gm;
  if (MODE == DENSE && p.comm_m_rot != 0) t.m_blk = (t.m_blk + p.comm_m_rot) % m_tiles;
  t.k_begin = 0;
  t.k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  if (MODE == DENSE && p.k_splits > 1) {
    const int per_split = (t.k_blocks + p.k_splits - 1) / p.k_splits;
    const int first = k_split * per_split;
    t.k_begin = first * BLOCK_K;
    t.k_blocks = max(0, min(per_split, t.k_blocks - first));
    t.valid = t.k_blocks > 0;  // an empty slice has nothing to add
  }
  t.group = 0;
  if (MODE == GROUPED_M) {
    t.group = p.tile_group[t.m_blk];
    t.valid = t.group >= 0;
  }
  return t;
}

__device__ __forceinline__ void comm_locate(const Params& p, int idx, int& owner, int& local) {
  const int blk = idx / p.comm_block_rows;
  owner = blk % p.comm_world;
  local = (blk / p.comm_world) * p.comm_block_rows + (idx - blk * p.comm_block_rows);
}

template <int MODE, int BLOCK_N, bool A_MN, bool B_MN, int EPI, int COMM = COMM_NONE, int CTAS = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_d, const Params p,
            const __grid_constant__ typename PeerArg<COMM>::type peers) {
  static_assert(COMM == COMM_NONE || MODE == DENSE, "fused communication is implemented for dense GEMMs");
  static_assert((COMM != COMM_AG_A && COMM != COMM_WAIT_A) || !A_MN, "COMM_AG_A / COMM_WAIT_A work on a K-major A");
  static_assert(COMM != COMM_AG_KA || A_MN, "COMM_AG_KA gathers an MN-major A along K");
  static_assert(COMM != COMM_AG_KB || B_MN, "COMM_AG_KB gathers an MN-major B along K");
  static_assert(COMM != COMM_RS_D || EPI == EPI_F32_ACC || EPI == EPI_BF16_ACC, "COMM_RS_D needs a reduce-add epilogue");
  static_assert(CTAS == 1 || (MODE == DENSE && COMM == COMM_NONE), "CTA pairs are implemented for plain dense GEMMs");
  using C = Cfg<BLOCK_N, CTAS>;
  constexpr int STAGES = C::STAGES;
  constexpr int TILE_M = C::TILE_M;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_STAGE_BYTES;
  uint8_t* smem_epi = smem + STAGES * C::STAGE_BYTES;  // 1024-aligned: STAGE_BYTES is a multiple of 8 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int total_tiles = num_tiles<MODE, BLOCK_N, TILE_M>(p);
  // persistent loop: one tile per CTA (pair) at a time
  const int pair_rank = (CTAS == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_tile = static_cast<int>(blockIdx.x) / CTAS;
  const int tile_stride = static_cast<int>(gridDim.x) / CTAS;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_epilogue) tma_prefetch_desc(&tmap_d);
    if constexpr (COMM != COMM_NONE && COMM != COMM_WAIT_A)
      for (int r = 0; r < p.comm_world; ++r) tma_prefetch_desc(&peers.m[r]);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS * CTAS);  // pair: the epilogue warps of both CTAs release the leader's accumulator
    }
    fence_barrier_init();
  }
  if constexpr (CTAS == 2) {
    if (warp == 1) tmem_alloc_pair<C::TMEM_COLS>(tmem_ptr_smem);  // executed by the same warp of both CTAs
    tc_fence_before();
    cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across the pair
    tc_fence_after();
  } else {
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // grouped modes read the tile's expert / row range from global memory: fetch it one whole tile ahead so that the
      // load latency never sits between two tiles on the critical path
      TileCoord t_next = get_tile<MODE, BLOCK_N, TILE_M>(p, first_tile);
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const TileCoord t = t_next;
        if (tile + tile_stride < total_tiles) t_next = get_tile<MODE, BLOCK_N, TILE_M>(p, tile + tile_stride);
        if (!t.valid) continue;
        // pair: this CTA's 128 rows of the 256-row tile and its half of the B tile
        const int m_idx = t.m_blk * TILE_M + pair_rank * BLOCK_M, n_idx = t.n_blk * BLOCK_N + pair_rank * C::B_ROWS;
        if constexpr (COMM == COMM_WAIT_A) {  // the shard holding this m-tile may still be in flight
          int owner, unused;
          comm_locate(p, m_idx, owner, unused);
          if (owner != p.comm_rank) {
            const int* flag = p.comm_flags + owner;
            int landed;
            long long spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(landed) : "l"(flag) : "memory");
              if (!landed) {
                __nanosleep(500);
                if (++spins > 8000000LL) {  // ~4 s: the producer of the flag never ran (would otherwise hang the GPU)
                  printf("d9d gemm: shard arrival flag %d never raised\n", owner);
                  __trap();
                }
              }
            } while (!landed);
            fence_proxy_async_all();  // the copy's generic-proxy writes are ordered before our async-proxy (TMA) reads
          }
        }
        for (int kb = 0; kb < t.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (CTAS == 2) {
            // both CTAs' loads complete on the leader's barrier (the MMA issuer waits there)
            if (pair_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CTAS * C::STAGE_BYTES);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          }
          const int k_idx = t.k_begin + kb * BLOCK_K;
          uint8_t* sa = smem_a + stage * C::A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * C::B_STAGE_BYTES;
          if constexpr (COMM == COMM_AG_A) {  // the m-tile lives in its owner's shard: load it over NVLink
            int owner, m_local;
            comm_locate(p, m_idx, owner, m_local);
            tma_load_3d(sa, &peers.m[owner], &full_bar[stage], k_idx, m_local, 0);
          } else if constexpr (COMM == COMM_AG_KA) {
            int owner, k_local;
            comm_locate(p, k_idx, owner, k_local);
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_3d(sa + j * (BLOCK_K * 128), &peers.m[owner], &full_bar[stage], m_idx + j * 64, k_local, 0);
          } else if (!A_MN) {
            if constexpr (CTAS == 2) tma_load_3d_pair(sa, &tmap_a, &full_bar[stage], k_idx, m_idx, 0);
            else tma_load_3d(sa, &tmap_a, &full_bar[stage], k_idx, m_idx, 0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) {
              if constexpr (CTAS == 2) tma_load_3d_pair(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m_idx + j * 64, k_idx, 0);
              else tma_load_3d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m_idx + j * 64, k_idx, 0);
            }
          }
          const int bg = (MODE == GROUPED_M) ? t.group : 0;
          if constexpr (COMM == COMM_AG_KB) {
            int owner, k_local;
            comm_locate(p, k_idx, owner, k_local);
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_3d(sb + j * (BLOCK_K * 128), &peers.m[owner], &full_bar[stage], n_idx + j * 64, k_local, 0);
          } else if (!B_MN) {
            if constexpr (CTAS == 2) tma_load_3d_pair(sb, &tmap_b, &full_bar[stage], k_idx, n_idx, bg);
            else tma_load_3d(sb, &tmap_b, &full_bar[stage], k_idx, n_idx, bg);
          } else {
#pragma unroll
            for (int j = 0; j < C::B_ROWS / 64; ++j) {
              if constexpr (CTAS == 2) tma_load_3d_pair(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n_idx + j * 64, k_idx, bg);
              else tma_load_3d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n_idx + j * 64, k_idx, bg);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (CTAS == 2) {
        // tail: every stage this CTA filled has been released again, i.e. all multicast arrivals of the leader's commits
        // have landed here before this CTA can leave the kernel
        for (int i = 0; i < STAGES; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair: the leader CTA issues for both) =================
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BLOCK_N, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    TileCoord t_next = get_tile<MODE, BLOCK_N, TILE_M>(p, first_tile);
    for (int tile = first_tile; tile < total_tiles && pair_rank == 0; tile += tile_stride) {
      const TileCoord t = t_next;
      if (tile + tile_stride < total_tiles) t_next = get_tile<MODE, BLOCK_N, TILE_M>(p, tile + tile_stride);
      if (!t.valid) continue;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < t.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(a_addr + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(b_addr + k * (UMMA_K * 2), 16, 1024);
            if constexpr (CTAS == 2) umma_f16_pair(tmem_d, da, db, idesc, (kb | k) != 0);
            else umma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          if constexpr (CTAS == 2) umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs
          else umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) {
        if constexpr (CTAS == 2) umma_commit_pair(&tmem_full[acc]);  // both CTAs' epilogues read their half of the accumulator
        else umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue (4 warps, one TMEM lane quadrant each) =================
    const int quad = warp & 3;
    const int lane = lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t epi_chunk = 0;  // running count of staged chunks (selects the staging buffer)
    TileCoord t_next = get_tile<MODE, BLOCK_N, TILE_M>(p, first_tile);
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
      const TileCoord t = t_next;
      if (tile + tile_stride < total_tiles) t_next = get_tile<MODE, BLOCK_N, TILE_M>(p, tile + tile_stride);
      if (!t.valid) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row0 = t.m_blk * TILE_M + pair_rank * BLOCK_M;  // first row of this CTA's 128-row slab
      const int row = row0 + quad * 32 + lane;
      const int col0 = t.n_blk * BLOCK_N;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
      const bool have_acc = t.k_blocks > 0;

      if constexpr (EPI == EPI_CE_LSE) {
        // online softmax statistics over this tile's columns
        float run_max = -INFINITY, run_sum = 0.f;
        long long tgt = row_ok ? p.ce_target[row] : -1;
        tgt = (tgt == p.ce_ignore_index) ? -1 : tgt - p.ce_col_offset;  // class id relative to this slice of the classifier
        const bool ce_xf = (p.ce_bias != nullptr) || (p.ce_softcap > 0.f);  // warp-uniform
        const float ce_inv_cap = p.ce_softcap > 0.f ? 1.f / p.ce_softcap : 0.f;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int cbase = col0 + c * 32;
          if (ce_xf) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float dz;
              const float b = (p.ce_bias != nullptr && cbase + i < p.N) ? p.ce_bias[cbase + i] : 0.f;
              r[i] = __float_as_uint(ce_transform(__uint_as_float(r[i]), b, p.ce_softcap, ce_inv_cap, dz));
            }
          }
          float cmax = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = (cbase + i < p.N) ? __uint_as_float(r[i]) : -INFINITY;
            r[i] = __float_as_uint(v);
            cmax = fmaxf(cmax, v);
          }
          if (cmax > -INFINITY) {
            const float new_max = fmaxf(run_max, cmax);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) s += __expf(__uint_as_float(r[i]) - new_max);
            run_sum = run_sum * __expf(run_max - new_max) + s;
            run_max = new_max;
          }
          if (row_ok && tgt >= cbase && tgt < cbase + 32 && tgt < p.N) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cbase + i == tgt) p.ce_tgt_logit[row] = __uint_as_float(r[i]);
          }
        }
        if (row_ok) {
          p.ce_part_max[static_cast<long long>(t.n_blk) * p.M + row] = run_max;
          p.ce_part_sum[static_cast<long long>(t.n_blk) * p.M + row] = run_sum;
        }
      } else if (p.tma_epilogue) {
        // ---- TMA epilogue: 128-byte output chunks (32 fp32 / 64 bf16 columns) staged in swizzled smem ----
        constexpr bool OUT_BF16 = (EPI == EPI_BF16 || EPI == EPI_BF16_ACC || EPI == EPI_CE_DLOGITS);
        constexpr bool REDUCE = (EPI == EPI_F32_ACC || EPI == EPI_BF16_ACC);
        constexpr int CHUNK_COLS = OUT_BF16 ? 64 : 32;
        static_assert(BLOCK_N % CHUNK_COLS == 0, "BLOCK_N must be a multiple of the epilogue chunk");
        float lse = 0.f, g = 0.f;
        long long tgt = -1;
        [[maybe_unused]] bool ce_xf = false;
        [[maybe_unused]] float ce_inv_cap = 0.f;
        if constexpr (EPI == EPI_CE_DLOGITS) {
          if (row_ok) {
            tgt = p.ce_target[row];
            lse = p.ce_lse[row];
            g = (tgt == p.ce_ignore_index) ? 0.f : p.ce_grad[row];
            tgt = (tgt == p.ce_ignore_index) ? -1 : tgt - p.ce_col_offset;
          }
          ce_xf = (p.ce_bias != nullptr) || (p.ce_softcap > 0.f);  // warp-uniform
          ce_inv_cap = p.ce_softcap > 0.f ? 1.f / p.ce_softcap : 0.f;
        }
        if (have_acc || !REDUCE) {
#pragma unroll 1
          for (int c = 0; c < BLOCK_N / CHUNK_COLS; ++c) {
            const int cbase = col0 + c * CHUNK_COLS;
            if (cbase >= p.N) break;  // warp-uniform
            uint8_t* stage_buf = smem_epi + (quad * C::EPI_BUFS + (epi_chunk % C::EPI_BUFS)) * 4096;
            ++epi_chunk;
            if (lane == 0) tma_store_wait_read<C::EPI_BUFS - 1>();  // the store that last read this buffer has drained
            __syncwarp();
            uint32_t packed[32];
            if constexpr (OUT_BF16) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                if (have_acc) {
                  tmem_ld_32x32b_x32(taddr + c * 64 + h * 32, r);
                  tmem_ld_wait();
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) r[i] = 0;
                }
                if constexpr (EPI == EPI_CE_DLOGITS) {
                  if (!ce_xf) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                      float pr = __expf(__uint_as_float(r[i]) - lse);
                      if (cbase + h * 32 + i == tgt) pr -= 1.f;
                      r[i] = __float_as_uint(pr * g);
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                      const int col = cbase + h * 32 + i;
                      float dz;
                      const float b = (p.ce_bias != nullptr && col < p.N) ? p.ce_bias[col] : 0.f;
                      float pr = __expf(ce_transform(__uint_as_float(r[i]), b, p.ce_softcap, ce_inv_cap, dz) - lse);
                      if (col == tgt) pr -= 1.f;
                      r[i] = __float_as_uint(pr * g * dz);
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  packed[h * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
              }
            } else {
              if (have_acc) {
                tmem_ld_32x32b_x32(taddr + c * 32, packed);
                tmem_ld_wait();
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) packed[i] = 0;
              }
            }
            // row `lane` of the 32 x 128B box; 16-byte chunk j lives at j ^ (row & 7) under the 128B swizzle
            uint8_t* my_row = stage_buf + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(my_row + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int r0 = row0 + quad * 32;
              const int g0 = (MODE == GROUPED_K) ? t.group : 0;
              if constexpr (COMM == COMM_RS_D) {  // reduce-add into the shard of the peer that owns these rows
                int owner, r_local;
                comm_locate(p, r0, owner, r_local);
                tma_reduce_add_3d(&peers.m[owner], stage_buf, cbase, r_local, 0);
              } else if constexpr (REDUCE) tma_reduce_add_3d(&tmap_d, stage_buf, cbase, r0, g0);
              else tma_store_3d(&tmap_d, stage_buf, cbase, r0, g0);
              tma_store_commit();
            }
          }
        }
      } else {
        long long d_off = static_cast<long long>(row) * p.ldd;
        if (MODE == GROUPED_K) d_off += static_cast<long long>(t.group) * p.d_group_stride;
        float lse = 0.f, g = 0.f;
        long long tgt = -1;
        [[maybe_unused]] bool ce_xf = false;
        [[maybe_unused]] float ce_inv_cap = 0.f;
        if constexpr (EPI == EPI_CE_DLOGITS) {
          if (row_ok) {
            tgt = p.ce_target[row];
            lse = p.ce_lse[row];
            g = (tgt == p.ce_ignore_index) ? 0.f : p.ce_grad[row];
            tgt = (tgt == p.ce_ignore_index) ? -1 : tgt - p.ce_col_offset;
          }
          ce_xf = (p.ce_bias != nullptr) || (p.ce_softcap > 0.f);  // warp-uniform
          ce_inv_cap = p.ce_softcap > 0.f ? 1.f / p.ce_softcap : 0.f;
        }
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          __syncwarp();
          if (have_acc) {
            tmem_ld_32x32b_x32(taddr + c * 32, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0;
          }
          const int cbase = col0 + c * 32;
          if (row_ok && cbase < p.N) {
          if constexpr (EPI == EPI_CE_DLOGITS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float dz = 1.f, z = __uint_as_float(r[i]);
              if (ce_xf) {
                const float b = (p.ce_bias != nullptr && cbase + i < p.N) ? p.ce_bias[cbase + i] : 0.f;
                z = ce_transform(z, b, p.ce_softcap, ce_inv_cap, dz);
              }
              float pr = __expf(z - lse);
              if (cbase + i == tgt) pr -= 1.f;
              r[i] = __float_as_uint(pr * g * dz);
            }
          }
          const bool full_chunk = (cbase + 32 <= p.N);
          if constexpr (EPI == EPI_BF16 || EPI == EPI_CE_DLOGITS || EPI == EPI_BF16_ACC) {
            __nv_bfloat16* dptr = reinterpret_cast<__nv_bfloat16*>(p.D) + d_off + cbase;
            const bool vec_ok = full_chunk && ((reinterpret_cast<uintptr_t>(dptr) & 15) == 0);
            if (vec_ok) {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                uint4 o;
                if constexpr (EPI == EPI_BF16_ACC) {
                  const uint4 old = *reinterpret_cast<const uint4*>(dptr + v * 8);
                  const float2 o0 = unpack_bf16x2(old.x), o1 = unpack_bf16x2(old.y), o2 = unpack_bf16x2(old.z),
                               o3 = unpack_bf16x2(old.w);
                  o.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]) + o0.x, __uint_as_float(r[v * 8 + 1]) + o0.y);
                  o.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]) + o1.x, __uint_as_float(r[v * 8 + 3]) + o1.y);
                  o.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]) + o2.x, __uint_as_float(r[v * 8 + 5]) + o2.y);
                  o.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]) + o3.x, __uint_as_float(r[v * 8 + 7]) + o3.y);
                } else {
                  o.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
                  o.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
                  o.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
                  o.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
                }
                *reinterpret_cast<uint4*>(dptr + v * 8) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (cbase + i < p.N) {
                  float v = __uint_as_float(r[i]);
                  if constexpr (EPI == EPI_BF16_ACC) v += __bfloat162float(dptr[i]);
                  dptr[i] = __float2bfloat16_rn(v);
                }
            }
          } else {
            float* dptr = reinterpret_cast<float*>(p.D) + d_off + cbase;
            const bool vec_ok = full_chunk && ((reinterpret_cast<uintptr_t>(dptr) & 15) == 0);
            if (vec_ok) {
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                float4 o = make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                                       __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
                if constexpr (EPI == EPI_F32_ACC) {
                  const float4 old = *reinterpret_cast<const float4*>(dptr + v * 4);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(dptr + v * 4) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (cbase + i < p.N) {
                  float v = __uint_as_float(r[i]);
                  if constexpr (EPI == EPI_F32_ACC) v += dptr[i];
                  dptr[i] = v;
                }
            }
          }
          }  // row_ok
        }
      }
      // release this accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_leader(&tmem_empty[acc]);
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();  // all bulk stores of this warp are complete before smem is released
  }

  tc_fence_before();
  if constexpr (CTAS == 2) {
    cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair may still touch its smem / barriers
    tc_fence_after();
    if (warp == 1) tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
  } else {
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace gemm
}  // namespace d9d
