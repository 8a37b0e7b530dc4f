// d9d_b200 native op launchers (plain C++ API, no torch dependency).
// Every launcher enqueues work on `stream` and returns immediately.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace d9d {

// ------------------------------------------------------------------ GEMM (tcgen05) ---------------
struct GemmArgs {
  int mode = 0;        // gemm::Mode
  int epi = 0;         // gemm::Epi
  bool a_mn = false;   // A stored [K, M] (M contiguous) instead of [M, K]
  bool b_mn = false;   // B stored [K, N] (N contiguous) instead of [N, K]
  const void* A = nullptr;
  const void* B = nullptr;
  void* D = nullptr;
  long long lda = 0, ldb = 0, ldd = 0;  // leading dims in elements
  long long b_group_stride = 0;         // GROUPED_M: elements between experts in B (0 => dense packing)
  long long d_group_stride = 0;         // GROUPED_K: elements between experts in D
  int M = 0, N = 0, K = 0;
  int num_groups = 1;
  long long k_total = 0;               // GROUPED_K: total rows of the grouped reduction dim
  const int* tile_group = nullptr;     // GROUPED_M
  const int* group_offsets = nullptr;  // GROUPED_K
  int block_n = 0;                     // 0 => heuristic
  int ctas = 0;                        // 1: one CTA per tile, 2: CTA pairs (cta_group::2, 256-row tiles), 0: process-wide policy
  int k_splits = 0;                    // dense fp32-accumulate epilogue only: 0 => heuristic, 1 => off
  // fused tensor-parallel communication (gemm::Comm); the sharded operand / output is described by its peers
  int comm = 0;
  int comm_world = 1;
  int comm_block_rows = 0;                     // rows per (rank, block) of the sharded dim
  const void* const* comm_peer_ptrs = nullptr; // host array [comm_world]: base pointer of every peer's shard
  long long comm_rows_local = 0;               // rows of one shard
  long long comm_ld = 0;                       // leading dim (elements) of the shards
  int comm_rank = 0;                           // COMM_WAIT_A: this rank
  const int* comm_flags = nullptr;             // COMM_WAIT_A: device array [comm_world] of shard arrival flags
  int comm_spare_sms = 0;                      // COMM_WAIT_A: SMs left free for the concurrent copy kernels
  // fused linear cross entropy
  const long long* ce_target = nullptr;
  const float* ce_lse = nullptr;
  const float* ce_grad = nullptr;
  float* ce_part_max = nullptr;
  float* ce_part_sum = nullptr;
  float* ce_tgt_logit = nullptr;
  long long ce_ignore_index = -100;
  float ce_softcap = 0.f;              // logits -> softcap * tanh(logits / softcap); 0 = off
  const float* ce_bias = nullptr;      // [N] fp32 logit bias
  long long ce_col_offset = 0;         // B is a row slice of the classifier starting at this class id
};

void gemm_dense(const GemmArgs& a, cudaStream_t stream);
void gemm_dense_pair(const GemmArgs& a, cudaStream_t stream);  // CTA-pair kernels (gemm_dense_pair.cu)
// process-wide policy for GemmArgs::ctas == 0: 0 = single-CTA kernels, 1 = CTA pairs for eligible dense / fused-CE shapes.
// Default 1 (D9D_GEMM_PAIR=0 switches it off); returns the previous value.
int gemm_set_pair_mode(int mode);
int gemm_pair_mode();
bool gemm_pair_eligible(const GemmArgs& a);
void gemm_grouped(const GemmArgs& a, cudaStream_t stream);
void gemm_ce(const GemmArgs& a, cudaStream_t stream);
void gemm_comm(const GemmArgs& a, cudaStream_t stream);  // dense GEMM with a.comm != 0 (gemm_comm_*.cu)
int gemm_ce_block_n();
// (part_max, part_sum)[n_tiles, M] + tgt_logit[M] -> lse[M], nll[M] (0 where target == ignore_index)
void ce_finalize(const float* part_max, const float* part_sum, const float* tgt_logit, const long long* target,
                 long long ignore_index, int n_tiles, int M, float* lse, float* nll, cudaStream_t stream);

// ------------------------------------------------------------------ fp8 GEMM (tcgen05 kind::f8f6f4 / mxf8f6f4) --
struct Fp8GemmArgs {
  const void* A = nullptr;   // [M, K] e4m3, K contiguous (row pitch lda bytes)
  const void* B = nullptr;   // [N, K] e4m3
  void* D = nullptr;         // [M, N] bf16
  long long lda = 0, ldb = 0, ldd = 0;
  int M = 0, N = 0, K = 0;
  bool block_scaled = false;          // true: MXFP8 (UE8M0 scale per 32 K-elements), false: row / column scales in the epilogue
  const float* scale_a = nullptr;     // [M] or nullptr
  const float* scale_b = nullptr;     // [N] or nullptr
  float scale_scalar = 1.f;
  const uint8_t* sfa = nullptr;       // [ceil(M/128), K/128, 512]
  const uint8_t* sfb = nullptr;       // [2*ceil(N/256), K/128, 512]
};
void gemm_fp8(const Fp8GemmArgs& a, cudaStream_t stream);
void quantize_rowwise_e4m3(const void* x, void* q, float* scale, long long M, int K, long long ldx, cudaStream_t stream);
// qt[C, R] (row pitch ldq bytes) = e4m3(x[R, C]^T / scale[c]); amax_scratch: [C] fp32
void quantize_colwise_t_e4m3(const void* x, void* qt, float* scale, float* amax_scratch, long long R, int C, long long ldx,
                             long long ldq, cudaStream_t stream);
// sf must be zero-initialised by the caller when M is not a multiple of 256 (pad blocks)
void quantize_mx_e4m3(const void* x, void* q, void* sf, long long M, int K, long long ldx, cudaStream_t stream);

// ------------------------------------------------------------------ normalisation ----------------
// dtype codes: 0 = bf16, 1 = fp32, 2 = fp16
void rms_norm_fwd(const void* x, const void* w, void* out, float* inv_rms, long long M, int N, float eps,
                  bool zero_centered, int dtype, cudaStream_t stream);
// dw_partial: [rms_norm_bwd_num_partials(), N] fp32 scratch; dw: [N] in dtype
int rms_norm_bwd_num_partials();
void rms_norm_bwd(const void* dout, const void* x, const void* w, const float* inv_rms, void* dx, void* dw,
                  float* dw_partial, long long M, int N, bool zero_centered, int dtype, cudaStream_t stream);

// ------------------------------------------------------------------ activations ------------------
void silu_mul_fwd(const void* x, const void* y, void* out, long long n, int dtype, cudaStream_t stream);
void silu_mul_bwd(const void* dout, const void* x, const void* y, void* dx, void* dy, long long n, int dtype,
                  cudaStream_t stream);
// MoE variant: out[r, :] = silu(x) * y * probs[r];  bwd also yields dprobs[r]
void silu_mul_probs_fwd(const void* x, const void* y, const float* probs, void* out, long long rows, int cols,
                        const int* valid_rows, cudaStream_t s);
// valid_rows (device scalar or nullptr): rows >= *valid_rows are skipped (unused tail of a static-capacity buffer)
void silu_mul_probs_bwd(const void* dout, const void* x, const void* y, const float* probs, void* dx, void* dy,
                        float* dprobs, long long rows, int cols, const int* valid_rows, cudaStream_t s);

// ------------------------------------------------------------------ stochastic rounding ----------
void sr_copy_f32_to_bf16(const float* src, void* dst, long long n, uint64_t seed, cudaStream_t stream);

struct AdamTensorMeta {
  void* p;        // bf16
  const void* g;  // bf16 or fp32
  void* m;        // bf16 or fp32
  void* v;        // bf16 or fp32
  long long n;
  long long rng_base;  // element offset into the launch-wide philox stream
};
// metas / block_map live in device memory. block_map[b] = {tensor index, chunk index}; chunk = ADAM_CHUNK elements.
constexpr int ADAM_CHUNK = 8192;
void adamw_sr_multi(const AdamTensorMeta* metas, const int2* block_map, int num_blocks, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, uint64_t seed,
                    const float* grad_scale /*device scalar or null*/, bool grad_bf16, bool state_bf16,
                    cudaStream_t stream);

// ------------------------------------------------------------------ RoPE (+ fused q/k RMSNorm) ---
// x: [T, H, D] (row stride = ldx elements per token), cos/sin cache: [max_pos, rope_dim] fp32 or bf16 (as float here),
// pos: [T] int64. style 0 = HALF, 1 = INTERLEAVED. inverse => apply the transposed rotation (backward).
void rope_apply(const void* x, void* out, const float* cos_cache, const float* sin_cache, const long long* pos,
                long long T, int H, int D, int rope_dim, long long ldx, long long ldo, int style, bool inverse,
                cudaStream_t stream);

// ------------------------------------------------------------------ MoE routing / permutation ----
// Builds the 128-row-aligned expert-sorted layout fully on device (no host sync):
//   topk_ids [T, k] int64 (ids outside [0, E) are dropped)
//   -> counts[E], seg_offsets[E+1] (aligned to `align`), row_map[T*k] (dest row or -1),
//      tile_group[cap/128] (expert per 128-row tile or -1), total_rows[1]
long long moe_layout_scratch_ints(long long n_entries, int E);
void moe_build_layout(const long long* topk_ids, long long T, int k, int E, int align, long long capacity_rows,
                      int* counts, int* seg_offsets, int* row_map, int* tile_group, int* chunk_scratch,
                      cudaStream_t stream);
// zero the pad rows [seg_offsets[e] + counts[e], seg_offsets[e+1]) of xp (and pp if given)
void moe_zero_pad(void* xp, float* pp, const int* counts, const int* seg_offsets, int E, int H, cudaStream_t stream);
// x[T,H] -> xp[cap,H] (rows not written stay zero: caller zero-fills), probs[T,k] -> pp[cap]
void moe_permute(const void* x, const float* probs, const int* row_map, void* xp, float* pp, long long T, int k,
                 int H, cudaStream_t stream);
// y[T,H] = sum_k yp[row_map[t,k], :]   (fp32 accumulate)
void moe_unpermute(const void* yp, const int* row_map, void* y, long long T, int k, int H, cudaStream_t stream);
// backward of permute: dx[T,H] = sum_k dxp[row_map[t,k]], dprobs[T,k] = dpp[row_map[t,k]]
void moe_permute_bwd(const void* dxp, const float* dpp, const int* row_map, void* dx, float* dprobs, long long T,
                     int k, int H, cudaStream_t stream);

// ------------------------------------------------------------------ grad utilities ---------------
// out[0] += sum(x^2) over a flat fp32/bf16 buffer (dtype code as above)
void sumsq_accumulate(const void* x, long long n, int dtype, float* out, cudaStream_t stream);
// x *= *scale (device scalar)
void scale_inplace(void* x, long long n, int dtype, const float* scale, cudaStream_t stream);

// ------------------------------------------------------------------ flash attention (forward / backward) --
// q / out / dout / dq [total_q, Hq, D], k / v / dk / dv [total_k, Hk, D] bf16 contiguous, where total = B * S for
// fixed-length batches (cu_q == cu_k == nullptr) or the packed token count (cu_q / cu_k: [B + 1] int32 device arrays, Sq /
// Sk = maximum sequence lengths).  lse / delta: fp32 [B, Hq, Sq] (packed: [Hq, total_q]).
// window: key k is visible to query q iff q + (Sk - Sq) - window_left <= k <= q + (Sk - Sq) + window_right (-1 = unbounded;
// causal = window_right 0).  softcap 0 = off.  sink: per-head logit joining the softmax denominator (or nullptr).
struct FlashAttnArgs {
  const void *q, *k, *v;
  void* out;            // forward: written; backward: read
  float* lse;           // forward: written; backward: read
  const void* dout;     // backward
  float* delta;         // backward scratch (written)
  void *dq, *dk, *dv;   // backward outputs
  const float* sink;
  const int *cu_q, *cu_k;
  int B, Sq, Sk, Hq, Hk, D;
  long long total_q, total_k;
  int window_left, window_right;
  float scale, softcap;
};
// variant: reserved (the forward hands P to the PV GEMM through swizzled shared memory; the backward kernels read P^T / dS^T
// as the A operand straight from tensor memory)
void flash_attn_fwd(const FlashAttnArgs& a, int variant, cudaStream_t stream);
// backward = delta pre-pass (delta[b,h,q] = sum_d dO*O, written to a.delta) followed by the dK/dV and dQ kernels (which read
// a.delta; the caller may subtract d(lse) from it in between)
void flash_attn_bwd_delta(const FlashAttnArgs& a, cudaStream_t stream);
void flash_attn_bwd(const FlashAttnArgs& a, cudaStream_t stream);

// ------------------------------------------------------------------ fused q/k RMSNorm + RoPE -------------
// q [T, Hq, D] (row stride ldq), k [T, Hk, D] (row stride ldk), cos/sin [T, rope_dim] fp32 (already gathered per token)
int qk_norm_rope_grid(long long T, int Ht);  // number of blocks == rows of dw_partial
void qk_norm_rope_fwd(const void* q, const void* k, const void* wq, const void* wk, const float* cos_t, const float* sin_t,
                      long long T, int Hq, int Hk, int D, int rope_dim, long long ldq, long long ldk, float eps,
                      bool zero_centered, int style, void* q_out, void* k_out, float* inv_rms, cudaStream_t s);
void qk_norm_rope_bwd(const void* dq_out, const void* dk_out, const void* q, const void* k, const void* wq, const void* wk,
                      const float* cos_t, const float* sin_t, const float* inv_rms, long long T, int Hq, int Hk, int D,
                      int rope_dim, long long ldq, long long ldk, bool zero_centered, int style, void* dq, void* dk,
                      float* dw_partial, float* dwq, float* dwk, cudaStream_t s);

// ------------------------------------------------------------------ MoE router -----------------------------
// logits [T, E] bf16 -> softmax (fp32) -> top-k by (prob + bias) -> idx [T, k] int64, probs [T, k] fp32 (renormalised)
void router_topk_fwd(const void* logits, const float* bias, long long T, int E, int k, bool renorm, long long* idx,
                     float* probs, cudaStream_t stream);
void router_topk_bwd(const void* logits, const long long* idx, const float* dprobs, long long T, int E, int k,
                     bool renorm, void* dlogits, cudaStream_t stream);

// ------------------------------------------------------------------ expert-parallel exchange over NVLink ----
// peer_base: device array [world] of symmetric-arena base pointers; off_* are byte offsets of regions inside the arena.
// dest_rank / dest_row: per (token, slot) pair, the owner rank and the row in its buffer (row < 0: dropped pair).
void ep_push(const void* x, const float* probs, const int* dest_rank, const int* dest_row, void* const* peer_base, long long off_x,
             long long off_p, long long n_pairs, int k, int H, cudaStream_t stream);
void ep_pull_sum(void* const* peer_base, long long off_y, long long off_dp, const int* dest_rank, const int* dest_row, void* y,
                 float* dprobs, long long T, int k, int H, cudaStream_t stream);

// ------------------------------------------------------------------ NVLink data-parallel optimizer -----------
// peer_* : device array [world] of per-replica base pointers of the symmetric arena; mc_* : multicast (NVLS) mapping
// of the same arena or nullptr.  [begin, end) is this rank's shard (multiples of 8 elements).
void nvl_reduce_shard(const float* const* peer_grads, const float* mc_grad, float* own_grad, long long begin,
                      long long end, int world, int rank, float* sumsq, cudaStream_t stream);
void nvl_adamw_shard(void* const* peer_params, void* mc_param, void* own_param, const float* own_grad, void* exp_avg,
                     void* exp_avg_sq, long long begin, long long end, int world, int rank, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float bias_corr1, float bias_corr2, uint64_t seed,
                     const float* grad_scale, bool state_bf16, cudaStream_t stream);

}  // namespace d9d
