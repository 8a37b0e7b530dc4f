// d9d_b200 — torch bindings (TORCH_LIBRARY) for the native sm_100a kernels.
// Compiled with the host compiler only; the kernels live in *.cu translation units that do not include torch.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/torch.h>

#include <tuple>
#include <vector>

#include "d9d_ops.h"

namespace {

using at::Tensor;

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

inline int dtype_code(const Tensor& t) {
  if (t.scalar_type() == at::kBFloat16) return 0;
  if (t.scalar_type() == at::kFloat) return 1;
  if (t.scalar_type() == at::kHalf) return 2;
  TORCH_CHECK(false, "d9d_b200: unsupported dtype ", t.scalar_type());
}

#define CHECK_CUDA_CONTIG(t) TORCH_CHECK((t).is_cuda() && (t).is_contiguous(), #t " must be a contiguous CUDA tensor")

// ------------------------------------------------------------------ GEMM ------------------------
int epi_for(const Tensor& d, bool accumulate) {
  if (d.scalar_type() == at::kBFloat16) return accumulate ? 3 : 0;
  TORCH_CHECK(d.scalar_type() == at::kFloat, "gemm output must be bf16 or fp32");
  return accumulate ? 2 : 1;
}

// d[M,N] (+)= op(a) · op(b)^T.  a: [M,K] or (a_mn) [K,M];  b: [N,K] or (b_mn) [K,N].  Row strides may be padded.
void gemm(const Tensor& a, const Tensor& b, Tensor d, bool a_mn, bool b_mn, bool accumulate, int64_t ctas) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && d.is_cuda(), "gemm: CUDA tensors required");
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16, "gemm: bf16 operands required");
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && d.dim() == 2, "gemm: 2-D tensors required");
  TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1 && d.stride(1) == 1, "gemm: innermost stride must be 1");
  c10::cuda::CUDAGuard guard(a.device());
  d9d::GemmArgs g;
  g.mode = 0;
  g.a_mn = a_mn; g.b_mn = b_mn;
  g.M = static_cast<int>(a_mn ? a.size(1) : a.size(0));
  g.K = static_cast<int>(a_mn ? a.size(0) : a.size(1));
  g.N = static_cast<int>(b_mn ? b.size(1) : b.size(0));
  TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == g.K, "gemm: K mismatch");
  TORCH_CHECK(d.size(0) == g.M && d.size(1) == g.N, "gemm: output shape mismatch");
  g.A = a.data_ptr(); g.B = b.data_ptr(); g.D = d.data_ptr();
  g.lda = a.stride(0); g.ldb = b.stride(0); g.ldd = d.stride(0);
  g.epi = epi_for(d, accumulate);
  g.ctas = static_cast<int>(ctas);
  d9d::gemm_dense(g, cur_stream());
}

int64_t gemm_set_pair_mode(int64_t mode) { return d9d::gemm_set_pair_mode(static_cast<int>(mode)); }

// d[R,N] = a[R,K] · W[e(r)];  b: [E,N,K] (b_mn=false) or [E,K,N] (b_mn=true)
void gemm_grouped_m(const Tensor& a, const Tensor& b, Tensor d, const Tensor& tile_group, bool b_mn, bool accumulate) {
  CHECK_CUDA_CONTIG(a); CHECK_CUDA_CONTIG(b); CHECK_CUDA_CONTIG(d); CHECK_CUDA_CONTIG(tile_group);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && d.scalar_type() == at::kBFloat16);
  TORCH_CHECK(tile_group.scalar_type() == at::kInt && a.size(0) % 128 == 0 && tile_group.numel() == a.size(0) / 128);
  c10::cuda::CUDAGuard guard(a.device());
  d9d::GemmArgs g;
  g.mode = 1; g.epi = accumulate ? 3 : 0; g.a_mn = false; g.b_mn = b_mn;
  g.M = static_cast<int>(a.size(0)); g.K = static_cast<int>(a.size(1));
  g.N = static_cast<int>(b_mn ? b.size(2) : b.size(1));
  TORCH_CHECK((b_mn ? b.size(1) : b.size(2)) == g.K, "gemm_grouped_m: K mismatch");
  TORCH_CHECK(d.size(0) == g.M && d.size(1) == g.N);
  g.num_groups = static_cast<int>(b.size(0));
  g.A = a.data_ptr(); g.B = b.data_ptr(); g.D = d.data_ptr();
  g.lda = a.stride(0); g.ldb = b.stride(1); g.ldd = d.stride(0);
  g.b_group_stride = b.stride(0);
  g.tile_group = tile_group.data_ptr<int>();
  d9d::gemm_grouped(g, cur_stream());
}

// d[E,M,N] (+)= a[rows_e, M]^T · b[rows_e, N] over each expert's row range group_offsets[e]..[e+1]
void gemm_grouped_k(const Tensor& a, const Tensor& b, Tensor d, const Tensor& group_offsets, bool accumulate) {
  CHECK_CUDA_CONTIG(a); CHECK_CUDA_CONTIG(b); CHECK_CUDA_CONTIG(d); CHECK_CUDA_CONTIG(group_offsets);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16);
  TORCH_CHECK(group_offsets.scalar_type() == at::kInt && a.size(0) == b.size(0));
  c10::cuda::CUDAGuard guard(a.device());
  d9d::GemmArgs g;
  g.mode = 2; g.a_mn = true; g.b_mn = true;
  g.M = static_cast<int>(a.size(1)); g.N = static_cast<int>(b.size(1)); g.K = 0;
  g.k_total = a.size(0);
  g.num_groups = static_cast<int>(d.size(0));
  TORCH_CHECK(group_offsets.numel() == g.num_groups + 1 && d.size(1) == g.M && d.size(2) == g.N);
  g.A = a.data_ptr(); g.B = b.data_ptr(); g.D = d.data_ptr();
  g.lda = a.stride(0); g.ldb = b.stride(0); g.ldd = d.stride(1); g.d_group_stride = d.stride(0);
  g.group_offsets = group_offsets.data_ptr<int>();
  g.epi = epi_for(d, accumulate);
  d9d::gemm_grouped(g, cur_stream());
}

// fused linear-CE forward over a row slice w = C[col_offset : col_offset + V] of the classifier:
// returns (nll[T], lse[T], tgt_logit[T]) fp32; tgt_logit is 0 where the target lies outside the slice (nll is then lse)
std::tuple<Tensor, Tensor, Tensor> ce_forward_ex(const Tensor& h, const Tensor& w, const Tensor& target, int64_t ignore_index,
                                                 const c10::optional<Tensor>& bias, double softcap, int64_t col_offset) {
  CHECK_CUDA_CONTIG(h); CHECK_CUDA_CONTIG(w); CHECK_CUDA_CONTIG(target);
  TORCH_CHECK(h.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && target.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard guard(h.device());
  const int T = static_cast<int>(h.size(0)), K = static_cast<int>(h.size(1)), V = static_cast<int>(w.size(0));
  TORCH_CHECK(w.size(1) == K && target.numel() == T);
  const int bn = d9d::gemm_ce_block_n();
  const int n_tiles = (V + bn - 1) / bn;
  auto fopt = h.options().dtype(at::kFloat);
  Tensor part_max = at::empty({n_tiles, T}, fopt), part_sum = at::empty({n_tiles, T}, fopt);
  Tensor tgt_logit = at::zeros({T}, fopt), lse = at::empty({T}, fopt), nll = at::empty({T}, fopt);
  d9d::GemmArgs g;
  g.mode = 0; g.epi = 4;
  g.M = T; g.N = V; g.K = K;
  g.A = h.data_ptr(); g.B = w.data_ptr(); g.lda = K; g.ldb = K; g.ldd = V;
  g.ce_target = reinterpret_cast<const long long*>(target.data_ptr<int64_t>());
  g.ce_part_max = part_max.data_ptr<float>(); g.ce_part_sum = part_sum.data_ptr<float>();
  g.ce_tgt_logit = tgt_logit.data_ptr<float>();
  g.ce_ignore_index = ignore_index;
  g.ce_softcap = static_cast<float>(softcap);
  g.ce_col_offset = col_offset;
  if (bias.has_value()) {
    CHECK_CUDA_CONTIG(*bias);
    TORCH_CHECK(bias->scalar_type() == at::kFloat && bias->numel() == V, "ce: bias must be fp32 [V]");
    g.ce_bias = bias->data_ptr<float>();
  }
  d9d::gemm_ce(g, cur_stream());
  d9d::ce_finalize(part_max.data_ptr<float>(), part_sum.data_ptr<float>(), tgt_logit.data_ptr<float>(),
                   reinterpret_cast<const long long*>(target.data_ptr<int64_t>()), ignore_index, n_tiles, T,
                   lse.data_ptr<float>(), nll.data_ptr<float>(), cur_stream());
  return {nll, lse, tgt_logit};
}

// fused linear-CE forward: returns (nll[T] fp32, lse[T] fp32)
std::tuple<Tensor, Tensor> ce_forward(const Tensor& h, const Tensor& w, const Tensor& target, int64_t ignore_index) {
  auto r = ce_forward_ex(h, w, target, ignore_index, c10::nullopt, 0.0, 0);
  return {std::get<0>(r), std::get<1>(r)};
}

// fused linear-CE backward, one chunk: out[T, V] (bf16) = grad[t] * d nll[t] / d logits[t, col_offset : col_offset + V]
// (w is the matching row slice of the classifier; lse is the log-sum-exp over the WHOLE vocabulary)
void ce_dlogits(const Tensor& h, const Tensor& w, const Tensor& target, const Tensor& lse, const Tensor& grad,
                Tensor out, int64_t ignore_index, const c10::optional<Tensor>& bias, double softcap, int64_t col_offset) {
  CHECK_CUDA_CONTIG(h); CHECK_CUDA_CONTIG(w); CHECK_CUDA_CONTIG(target); CHECK_CUDA_CONTIG(lse);
  CHECK_CUDA_CONTIG(grad);
  TORCH_CHECK(out.is_cuda() && out.dim() == 2 && out.stride(1) == 1, "ce_dlogits: out must be row-major (row pitch may be padded)");
  TORCH_CHECK(out.scalar_type() == at::kBFloat16 && lse.scalar_type() == at::kFloat && grad.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(h.device());
  const int T = static_cast<int>(h.size(0)), K = static_cast<int>(h.size(1)), V = static_cast<int>(w.size(0));
  TORCH_CHECK(out.size(0) == T && out.size(1) == V);
  d9d::GemmArgs g;
  g.mode = 0; g.epi = 5;
  g.M = T; g.N = V; g.K = K;
  g.A = h.data_ptr(); g.B = w.data_ptr(); g.D = out.data_ptr(); g.lda = K; g.ldb = K; g.ldd = out.stride(0);
  g.ce_target = reinterpret_cast<const long long*>(target.data_ptr<int64_t>());
  g.ce_lse = lse.data_ptr<float>(); g.ce_grad = grad.data_ptr<float>();
  g.ce_ignore_index = ignore_index;
  g.ce_softcap = static_cast<float>(softcap);
  g.ce_col_offset = col_offset;
  if (bias.has_value()) {
    CHECK_CUDA_CONTIG(*bias);
    TORCH_CHECK(bias->scalar_type() == at::kFloat && bias->numel() == V, "ce: bias must be fp32 [V]");
    g.ce_bias = bias->data_ptr<float>();
  }
  d9d::gemm_ce(g, cur_stream());
}

// ------------------------------------------------------------------ fp8 GEMM -----------------------
// x[M,K] bf16 -> (q[M,K] e4m3, scale[M] fp32) with x ~= q * scale[:, None]
std::tuple<Tensor, Tensor> quantize_rowwise(const Tensor& x) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1 && x.scalar_type() == at::kBFloat16, "quantize_rowwise: 2-D bf16 CUDA tensor required");
  TORCH_CHECK(x.size(1) % 16 == 0 && x.stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15) == 0,
              "quantize_rowwise: K must be a multiple of 16 and rows 16-byte aligned");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t M = x.size(0), K = x.size(1);
  Tensor q = at::empty({M, K}, x.options().dtype(at::kFloat8_e4m3fn));
  Tensor scale = at::empty({M}, x.options().dtype(at::kFloat));
  d9d::quantize_rowwise_e4m3(x.data_ptr(), q.data_ptr(), scale.data_ptr<float>(), M, static_cast<int>(K), x.stride(0), cur_stream());
  return {q, scale};
}

// x[R,C] bf16 -> (qt[C,R] e4m3 (row pitch padded to 16 bytes), scale[C] fp32) with x[r,c] ~= qt[c,r] * scale[c]
std::tuple<Tensor, Tensor> quantize_colwise_t(const Tensor& x) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1 && x.scalar_type() == at::kBFloat16, "quantize_colwise_t: 2-D bf16 CUDA tensor required");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t R = x.size(0), C = x.size(1);
  const int64_t pitch = (R + 15) / 16 * 16;
  Tensor qt = at::zeros({C, pitch}, x.options().dtype(at::kFloat8_e4m3fn));
  Tensor scale = at::empty({C}, x.options().dtype(at::kFloat));
  Tensor amax = at::empty({C}, x.options().dtype(at::kFloat));
  d9d::quantize_colwise_t_e4m3(x.data_ptr(), qt.data_ptr(), scale.data_ptr<float>(), amax.data_ptr<float>(), R, static_cast<int>(C),
                               x.stride(0), pitch, cur_stream());
  return {qt.narrow(1, 0, R), scale};
}

// x[M,K] bf16 -> (q[M,K] e4m3, sf uint8 [2*ceil(M/256), K/128, 512]): OCP MXFP8, one UE8M0 scale per 32 elements along K,
// stored in the 128-row x 4-group block layout tcgen05.cp expects
std::tuple<Tensor, Tensor> quantize_mx(const Tensor& x) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1 && x.scalar_type() == at::kBFloat16, "quantize_mx: 2-D bf16 CUDA tensor required");
  TORCH_CHECK(x.size(1) % 128 == 0 && x.stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15) == 0,
              "quantize_mx: K must be a multiple of 128 and rows 16-byte aligned");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t M = x.size(0), K = x.size(1);
  Tensor q = at::empty({M, K}, x.options().dtype(at::kFloat8_e4m3fn));
  Tensor sf = at::zeros({(M + 255) / 256 * 2, K / 128, 512}, x.options().dtype(at::kByte));
  d9d::quantize_mx_e4m3(x.data_ptr(), q.data_ptr(), sf.data_ptr(), M, static_cast<int>(K), x.stride(0), cur_stream());
  return {q, sf};
}

static void fill_fp8_operands(d9d::Fp8GemmArgs& g, const Tensor& a, const Tensor& b, Tensor& d) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && d.is_cuda() && a.dim() == 2 && b.dim() == 2 && d.dim() == 2, "gemm_fp8: 2-D CUDA tensors required");
  TORCH_CHECK(a.scalar_type() == at::kFloat8_e4m3fn && b.scalar_type() == at::kFloat8_e4m3fn, "gemm_fp8: e4m3 operands required");
  TORCH_CHECK(d.scalar_type() == at::kBFloat16 && a.stride(1) == 1 && b.stride(1) == 1 && d.stride(1) == 1, "gemm_fp8: bf16 row-major output");
  TORCH_CHECK(a.size(1) == b.size(1) && d.size(0) == a.size(0) && d.size(1) == b.size(0), "gemm_fp8: shape mismatch");
  g.M = static_cast<int>(a.size(0)); g.K = static_cast<int>(a.size(1)); g.N = static_cast<int>(b.size(0));
  g.A = a.data_ptr(); g.B = b.data_ptr(); g.D = d.data_ptr();
  g.lda = a.stride(0); g.ldb = b.stride(0); g.ldd = d.stride(0);
}

// d[M,N] bf16 = (a[M,K] · b[N,K]^T) * scale_a[m] * scale_b[n] * scalar
void gemm_fp8(const Tensor& a, const Tensor& b, const c10::optional<Tensor>& scale_a, const c10::optional<Tensor>& scale_b, double scalar,
              Tensor d) {
  c10::cuda::CUDAGuard guard(a.device());
  d9d::Fp8GemmArgs g;
  fill_fp8_operands(g, a, b, d);
  if (scale_a.has_value()) {
    CHECK_CUDA_CONTIG(*scale_a);
    TORCH_CHECK(scale_a->scalar_type() == at::kFloat && scale_a->numel() == g.M, "gemm_fp8: scale_a must be fp32 [M]");
    g.scale_a = scale_a->data_ptr<float>();
  }
  if (scale_b.has_value()) {
    CHECK_CUDA_CONTIG(*scale_b);
    TORCH_CHECK(scale_b->scalar_type() == at::kFloat && scale_b->numel() == g.N, "gemm_fp8: scale_b must be fp32 [N]");
    g.scale_b = scale_b->data_ptr<float>();
  }
  g.scale_scalar = static_cast<float>(scalar);
  d9d::gemm_fp8(g, cur_stream());
}

// d[M,N] bf16 = sum_k (a * 2^sfa) (b * 2^sfb): block-scaled (MXFP8) operands from quantize_mx
void gemm_mxfp8(const Tensor& a, const Tensor& sfa, const Tensor& b, const Tensor& sfb, Tensor d) {
  c10::cuda::CUDAGuard guard(a.device());
  d9d::Fp8GemmArgs g;
  fill_fp8_operands(g, a, b, d);
  CHECK_CUDA_CONTIG(sfa); CHECK_CUDA_CONTIG(sfb);
  TORCH_CHECK(g.K % 128 == 0, "gemm_mxfp8: K must be a multiple of 128");
  const int64_t kb = g.K / 128;
  TORCH_CHECK(sfa.scalar_type() == at::kByte && sfa.numel() >= (static_cast<int64_t>(g.M) + 127) / 128 * kb * 512, "gemm_mxfp8: sfa too small");
  TORCH_CHECK(sfb.scalar_type() == at::kByte && sfb.numel() >= (static_cast<int64_t>(g.N) + 255) / 256 * 2 * kb * 512, "gemm_mxfp8: sfb too small");
  g.block_scaled = true;
  g.sfa = sfa.data_ptr<uint8_t>(); g.sfb = sfb.data_ptr<uint8_t>();
  d9d::gemm_fp8(g, cur_stream());
}

// ------------------------------------------------------------------ RMSNorm ---------------------
std::tuple<Tensor, Tensor> rms_norm_fwd(const Tensor& x, const Tensor& w, double eps, bool zero_centered) {
  CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(w);
  TORCH_CHECK(x.scalar_type() == w.scalar_type(), "rms_norm: x and weight dtypes must match");
  c10::cuda::CUDAGuard guard(x.device());
  const int N = static_cast<int>(x.size(-1));
  const int64_t M = x.numel() / N;
  Tensor out = at::empty_like(x);
  Tensor inv = at::empty({M}, x.options().dtype(at::kFloat));
  d9d::rms_norm_fwd(x.data_ptr(), w.data_ptr(), out.data_ptr(), inv.data_ptr<float>(), M, N, static_cast<float>(eps),
                    zero_centered, dtype_code(x), cur_stream());
  return {out, inv};
}

std::tuple<Tensor, Tensor> rms_norm_bwd(const Tensor& dout, const Tensor& x, const Tensor& w, const Tensor& inv_rms,
                                        bool zero_centered) {
  CHECK_CUDA_CONTIG(dout); CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(w); CHECK_CUDA_CONTIG(inv_rms);
  c10::cuda::CUDAGuard guard(x.device());
  const int N = static_cast<int>(x.size(-1));
  const int64_t M = x.numel() / N;
  Tensor dx = at::empty_like(x);
  Tensor dw = at::empty_like(w);
  Tensor partial = at::empty({d9d::rms_norm_bwd_num_partials(), N}, x.options().dtype(at::kFloat));
  d9d::rms_norm_bwd(dout.data_ptr(), x.data_ptr(), w.data_ptr(), inv_rms.data_ptr<float>(), dx.data_ptr(),
                    dw.data_ptr(), partial.data_ptr<float>(), M, N, zero_centered, dtype_code(x), cur_stream());
  return {dx, dw};
}

// ------------------------------------------------------------------ SiLU * mul -------------------
Tensor silu_mul_fwd(const Tensor& x, const Tensor& y) {
  CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(y);
  TORCH_CHECK(x.scalar_type() == y.scalar_type() && x.numel() == y.numel());
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty_like(x);
  d9d::silu_mul_fwd(x.data_ptr(), y.data_ptr(), out.data_ptr(), x.numel(), dtype_code(x), cur_stream());
  return out;
}

std::tuple<Tensor, Tensor> silu_mul_bwd(const Tensor& dout, const Tensor& x, const Tensor& y) {
  CHECK_CUDA_CONTIG(dout); CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(y);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor dx = at::empty_like(x), dy = at::empty_like(y);
  d9d::silu_mul_bwd(dout.data_ptr(), x.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.data_ptr(), x.numel(),
                    dtype_code(x), cur_stream());
  return {dx, dy};
}

static const int* valid_rows_ptr(const c10::optional<Tensor>& v) {
  if (!v.has_value()) return nullptr;
  TORCH_CHECK(v->is_cuda() && v->scalar_type() == at::kInt && v->numel() >= 1, "valid_rows must be an int32 CUDA scalar");
  return v->data_ptr<int>();
}

Tensor silu_mul_probs_fwd(const Tensor& x, const Tensor& y, const Tensor& probs, const c10::optional<Tensor>& valid_rows) {
  CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(y); CHECK_CUDA_CONTIG(probs);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && y.scalar_type() == at::kBFloat16 && probs.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty_like(x);
  d9d::silu_mul_probs_fwd(x.data_ptr(), y.data_ptr(), probs.data_ptr<float>(), out.data_ptr(), x.size(0),
                          static_cast<int>(x.size(1)), valid_rows_ptr(valid_rows), cur_stream());
  return out;
}

std::tuple<Tensor, Tensor, Tensor> silu_mul_probs_bwd(const Tensor& dout, const Tensor& x, const Tensor& y,
                                                      const Tensor& probs, const c10::optional<Tensor>& valid_rows) {
  CHECK_CUDA_CONTIG(dout); CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(y); CHECK_CUDA_CONTIG(probs);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor dx = at::empty_like(x), dy = at::empty_like(y), dp = at::empty_like(probs);
  d9d::silu_mul_probs_bwd(dout.data_ptr(), x.data_ptr(), y.data_ptr(), probs.data_ptr<float>(), dx.data_ptr(),
                          dy.data_ptr(), dp.data_ptr<float>(), x.size(0), static_cast<int>(x.size(1)), valid_rows_ptr(valid_rows),
                          cur_stream());
  return {dx, dy, dp};
}

// ------------------------------------------------------------------ stochastic rounding ----------
void sr_copy_(Tensor dst, const Tensor& src, int64_t seed) {
  CHECK_CUDA_CONTIG(dst); CHECK_CUDA_CONTIG(src);
  TORCH_CHECK(dst.scalar_type() == at::kBFloat16 && src.scalar_type() == at::kFloat && dst.numel() == src.numel());
  c10::cuda::CUDAGuard guard(src.device());
  d9d::sr_copy_f32_to_bf16(src.data_ptr<float>(), dst.data_ptr(), src.numel(), static_cast<uint64_t>(seed), cur_stream());
}

// metas: int64 [n_tensors, 6] = (p, g, m, v, numel, rng_base) in device memory; block_map: int32 [n_blocks, 2]
void adamw_sr_multi_(const Tensor& metas, const Tensor& block_map, double lr, double beta1, double beta2, double eps,
                     double weight_decay, int64_t step, int64_t seed, const c10::optional<Tensor>& grad_scale,
                     bool grad_bf16, bool state_bf16) {
  CHECK_CUDA_CONTIG(metas); CHECK_CUDA_CONTIG(block_map);
  TORCH_CHECK(metas.scalar_type() == at::kLong && metas.size(1) == 6 && block_map.scalar_type() == at::kInt);
  static_assert(sizeof(d9d::AdamTensorMeta) == 6 * sizeof(int64_t), "meta layout");
  c10::cuda::CUDAGuard guard(metas.device());
  const double bc1 = 1.0 - std::exp(static_cast<double>(step) * std::log(beta1));
  const double bc2 = 1.0 - std::exp(static_cast<double>(step) * std::log(beta2));
  const float* gs = nullptr;
  if (grad_scale.has_value()) {
    TORCH_CHECK(grad_scale->is_cuda() && grad_scale->scalar_type() == at::kFloat && grad_scale->numel() == 1);
    gs = grad_scale->data_ptr<float>();
  }
  d9d::adamw_sr_multi(reinterpret_cast<const d9d::AdamTensorMeta*>(metas.data_ptr<int64_t>()),
                      reinterpret_cast<const int2*>(block_map.data_ptr<int>()), static_cast<int>(block_map.size(0)),
                      static_cast<float>(lr), static_cast<float>(beta1), static_cast<float>(beta2),
                      static_cast<float>(eps), static_cast<float>(weight_decay), static_cast<float>(bc1),
                      static_cast<float>(bc2), static_cast<uint64_t>(seed), gs, grad_bf16, state_bf16, cur_stream());
}

// ------------------------------------------------------------------ RoPE -------------------------
// x: [T, H, D] view (token stride arbitrary, head/dim contiguous); cos/sin: [max_pos, rope_dim] fp32; pos: [T] int64
Tensor rope_apply(const Tensor& x, const Tensor& cos_cache, const Tensor& sin_cache, const Tensor& pos, int64_t style,
                  bool inverse) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 3 && x.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x.stride(2) == 1 && x.stride(1) == x.size(2), "rope: heads must be contiguous");
  CHECK_CUDA_CONTIG(cos_cache); CHECK_CUDA_CONTIG(sin_cache); CHECK_CUDA_CONTIG(pos);
  TORCH_CHECK(cos_cache.scalar_type() == at::kFloat && sin_cache.scalar_type() == at::kFloat && pos.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t T = x.size(0);
  const int H = static_cast<int>(x.size(1)), D = static_cast<int>(x.size(2));
  Tensor out = at::empty({T, H, D}, x.options());
  d9d::rope_apply(x.data_ptr(), out.data_ptr(), cos_cache.data_ptr<float>(), sin_cache.data_ptr<float>(),
                  reinterpret_cast<const long long*>(pos.data_ptr<int64_t>()), T, H, D,
                  static_cast<int>(cos_cache.size(1)), x.stride(0), static_cast<int64_t>(H) * D,
                  static_cast<int>(style), inverse, cur_stream());
  return out;
}

// ------------------------------------------------------------------ MoE --------------------------
// -> (counts[E], seg_offsets[E+1], row_map[T*k], tile_group[capacity/128])   all int32, no host sync
std::tuple<Tensor, Tensor, Tensor, Tensor> moe_build_layout(const Tensor& topk_ids, int64_t num_experts, int64_t align,
                                                            int64_t capacity) {
  CHECK_CUDA_CONTIG(topk_ids);
  TORCH_CHECK(topk_ids.scalar_type() == at::kLong && topk_ids.dim() == 2);
  c10::cuda::CUDAGuard guard(topk_ids.device());
  const int64_t T = topk_ids.size(0), k = topk_ids.size(1);
  auto iopt = topk_ids.options().dtype(at::kInt);
  Tensor counts = at::empty({num_experts}, iopt), seg = at::empty({num_experts + 1}, iopt);
  Tensor row_map = at::empty({T * k}, iopt), tile_group = at::empty({capacity / 128}, iopt);
  Tensor scratch = at::empty({std::max<int64_t>(1, d9d::moe_layout_scratch_ints(T * k, static_cast<int>(num_experts)))}, iopt);
  d9d::moe_build_layout(reinterpret_cast<const long long*>(topk_ids.data_ptr<int64_t>()), T, static_cast<int>(k),
                        static_cast<int>(num_experts), static_cast<int>(align), capacity, counts.data_ptr<int>(),
                        seg.data_ptr<int>(), row_map.data_ptr<int>(), tile_group.data_ptr<int>(),
                        scratch.data_ptr<int>(), cur_stream());
  return {counts, seg, row_map, tile_group};
}

// x[T,H] (+ probs[T,k]) -> xp[capacity,H] (+ pp[capacity]); pad rows zeroed, rows past the last segment untouched
std::tuple<Tensor, Tensor> moe_permute(const Tensor& x, const c10::optional<Tensor>& probs, const Tensor& row_map,
                                       const Tensor& counts, const Tensor& seg_offsets, int64_t capacity) {
  CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(row_map);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && x.dim() == 2);
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t T = x.size(0), H = x.size(1);
  const int k = static_cast<int>(row_map.numel() / std::max<int64_t>(T, 1));
  Tensor xp = at::empty({capacity, H}, x.options());
  Tensor pp;
  const float* pptr = nullptr;
  float* ppout = nullptr;
  if (probs.has_value()) {
    TORCH_CHECK(probs->is_contiguous() && probs->scalar_type() == at::kFloat);
    pp = at::empty({capacity}, x.options().dtype(at::kFloat));
    pptr = probs->data_ptr<float>();
    ppout = pp.data_ptr<float>();
  } else {
    pp = at::empty({0}, x.options().dtype(at::kFloat));
  }
  d9d::moe_permute(x.data_ptr(), pptr, row_map.data_ptr<int>(), xp.data_ptr(), ppout, T, k, static_cast<int>(H), cur_stream());
  d9d::moe_zero_pad(xp.data_ptr(), ppout, counts.data_ptr<int>(), seg_offsets.data_ptr<int>(),
                    static_cast<int>(counts.numel()), static_cast<int>(H), cur_stream());
  return {xp, pp};
}

// zero the pad rows [seg[e] + counts[e], seg[e+1]) of an aligned expert-sorted buffer (and of its per-row fp32 side array)
void moe_zero_pad(Tensor xp, const c10::optional<Tensor>& pp, const Tensor& counts, const Tensor& seg_offsets) {
  TORCH_CHECK(xp.is_cuda() && xp.dim() == 2 && xp.scalar_type() == at::kBFloat16 && xp.stride(1) == 1 && xp.stride(0) == xp.size(1));
  TORCH_CHECK(counts.scalar_type() == at::kInt && seg_offsets.scalar_type() == at::kInt && counts.is_contiguous() && seg_offsets.is_contiguous());
  c10::cuda::CUDAGuard guard(xp.device());
  float* p = nullptr;
  if (pp.has_value()) {
    TORCH_CHECK(pp->scalar_type() == at::kFloat && pp->is_contiguous());
    p = pp->data_ptr<float>();
  }
  d9d::moe_zero_pad(xp.data_ptr(), p, counts.data_ptr<int>(), seg_offsets.data_ptr<int>(), static_cast<int>(counts.numel()),
                    static_cast<int>(xp.size(1)), cur_stream());
}

// y[T,H] = sum_j yp[row_map[t,j]];   with dpp: also dprobs[T,k] = dpp[row_map[t,j]]
std::tuple<Tensor, Tensor> moe_gather(const Tensor& yp, const c10::optional<Tensor>& dpp, const Tensor& row_map,
                                      int64_t T, int64_t k) {
  CHECK_CUDA_CONTIG(yp); CHECK_CUDA_CONTIG(row_map);
  TORCH_CHECK(yp.scalar_type() == at::kBFloat16 && row_map.numel() == T * k);
  c10::cuda::CUDAGuard guard(yp.device());
  const int H = static_cast<int>(yp.size(1));
  Tensor y = at::empty({T, H}, yp.options());
  Tensor dprobs;
  if (dpp.has_value()) {
    dprobs = at::empty({T, k}, yp.options().dtype(at::kFloat));
    d9d::moe_permute_bwd(yp.data_ptr(), dpp->data_ptr<float>(), row_map.data_ptr<int>(), y.data_ptr(),
                         dprobs.data_ptr<float>(), T, static_cast<int>(k), H, cur_stream());
  } else {
    dprobs = at::empty({0}, yp.options().dtype(at::kFloat));
    d9d::moe_unpermute(yp.data_ptr(), row_map.data_ptr<int>(), y.data_ptr(), T, static_cast<int>(k), H, cur_stream());
  }
  return {y, dprobs};
}

// ------------------------------------------------------------------ grad utils -------------------
void sumsq_accumulate_(const Tensor& x, Tensor out) {
  CHECK_CUDA_CONTIG(x);
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kFloat && out.numel() == 1);
  c10::cuda::CUDAGuard guard(x.device());
  d9d::sumsq_accumulate(x.data_ptr(), x.numel(), dtype_code(x), out.data_ptr<float>(), cur_stream());
}

void scale_inplace_(Tensor x, const Tensor& scale) {
  CHECK_CUDA_CONTIG(x);
  TORCH_CHECK(scale.is_cuda() && scale.scalar_type() == at::kFloat && scale.numel() == 1);
  c10::cuda::CUDAGuard guard(x.device());
  d9d::scale_inplace(x.data_ptr(), x.numel(), dtype_code(x), scale.data_ptr<float>(), cur_stream());
}

}  // namespace

// ------------------------------------------------------------------ flash attention forward / backward
// q [B, Sq, Hq, D] (or packed [total_q, Hq, D] with cu_seqlens), k / v likewise.  window_left / window_right: -1 = unbounded.
static d9d::FlashAttnArgs fa_args(const Tensor& q, const Tensor& k, const Tensor& v, double scale, int64_t window_left,
                                  int64_t window_right, double softcap, const c10::optional<Tensor>& sink,
                                  const c10::optional<Tensor>& cu_q, const c10::optional<Tensor>& cu_k, int64_t max_q, int64_t max_k) {
  CHECK_CUDA_CONTIG(q); CHECK_CUDA_CONTIG(k); CHECK_CUDA_CONTIG(v);
  TORCH_CHECK(q.scalar_type() == at::kBFloat16 && k.scalar_type() == at::kBFloat16 && v.scalar_type() == at::kBFloat16,
              "flash_attn: bf16 tensors expected");
  TORCH_CHECK(k.sizes() == v.sizes(), "flash_attn: k and v must have the same shape");
  d9d::FlashAttnArgs a{};
  a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr();
  const bool varlen = cu_q.has_value();
  TORCH_CHECK(varlen == cu_k.has_value(), "flash_attn: cu_seqlens_q and cu_seqlens_k go together");
  if (varlen) {
    TORCH_CHECK(q.dim() == 3 && k.dim() == 3, "flash_attn (varlen): [total, H, D] tensors expected");
    TORCH_CHECK(cu_q->is_cuda() && cu_k->is_cuda() && cu_q->scalar_type() == at::kInt && cu_k->scalar_type() == at::kInt &&
                cu_q->is_contiguous() && cu_k->is_contiguous() && cu_q->numel() == cu_k->numel() && cu_q->numel() >= 2,
                "flash_attn (varlen): cu_seqlens must be int32 CUDA tensors of length B + 1");
    a.B = static_cast<int>(cu_q->numel() - 1);
    a.Sq = static_cast<int>(max_q); a.Sk = static_cast<int>(max_k);
    a.total_q = q.size(0); a.total_k = k.size(0);
    a.Hq = static_cast<int>(q.size(1)); a.Hk = static_cast<int>(k.size(1)); a.D = static_cast<int>(q.size(2));
    TORCH_CHECK(k.size(2) == a.D);
    a.cu_q = cu_q->data_ptr<int>(); a.cu_k = cu_k->data_ptr<int>();
  } else {
    TORCH_CHECK(q.dim() == 4 && k.dim() == 4, "flash_attn: [B, S, H, D] tensors expected");
    TORCH_CHECK(q.size(0) == k.size(0) && q.size(3) == k.size(3));
    a.B = static_cast<int>(q.size(0)); a.Sq = static_cast<int>(q.size(1)); a.Sk = static_cast<int>(k.size(1));
    a.Hq = static_cast<int>(q.size(2)); a.Hk = static_cast<int>(k.size(2)); a.D = static_cast<int>(q.size(3));
    a.total_q = static_cast<long long>(a.B) * a.Sq; a.total_k = static_cast<long long>(a.B) * a.Sk;
  }
  a.window_left = static_cast<int>(window_left); a.window_right = static_cast<int>(window_right);
  a.scale = static_cast<float>(scale); a.softcap = static_cast<float>(softcap);
  if (sink.has_value()) {
    TORCH_CHECK(sink->is_cuda() && sink->scalar_type() == at::kFloat && sink->is_contiguous() && sink->numel() == a.Hq,
                "flash_attn: sink must be a contiguous fp32 [Hq] CUDA tensor");
    a.sink = sink->data_ptr<float>();
  }
  return a;
}

static Tensor fa_stats_like(const d9d::FlashAttnArgs& a, const Tensor& q) {
  return a.cu_q != nullptr ? at::empty({a.Hq, a.total_q}, q.options().dtype(at::kFloat))
                           : at::empty({a.B, a.Hq, a.Sq}, q.options().dtype(at::kFloat));
}

std::tuple<Tensor, Tensor> flash_attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, double scale, int64_t window_left,
                                          int64_t window_right, double softcap, const c10::optional<Tensor>& sink,
                                          const c10::optional<Tensor>& cu_q, const c10::optional<Tensor>& cu_k, int64_t max_q,
                                          int64_t max_k, int64_t variant) {
  c10::cuda::CUDAGuard guard(q.device());
  d9d::FlashAttnArgs a = fa_args(q, k, v, scale, window_left, window_right, softcap, sink, cu_q, cu_k, max_q, max_k);
  Tensor out = at::empty_like(q);
  Tensor lse = fa_stats_like(a, q);
  a.out = out.data_ptr();
  a.lse = lse.data_ptr<float>();
  d9d::flash_attn_fwd(a, static_cast<int>(variant), cur_stream());
  return {out, lse};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> flash_attn_bwd(const Tensor& dout, const Tensor& q, const Tensor& k, const Tensor& v,
                                                          const Tensor& out, const Tensor& lse, double scale, int64_t window_left,
                                                          int64_t window_right, double softcap, const c10::optional<Tensor>& cu_q,
                                                          const c10::optional<Tensor>& cu_k, int64_t max_q, int64_t max_k,
                                                          const c10::optional<Tensor>& dlse) {
  c10::cuda::CUDAGuard guard(q.device());
  CHECK_CUDA_CONTIG(dout); CHECK_CUDA_CONTIG(out); CHECK_CUDA_CONTIG(lse);
  TORCH_CHECK(dout.scalar_type() == at::kBFloat16 && out.scalar_type() == at::kBFloat16 && lse.scalar_type() == at::kFloat);
  TORCH_CHECK(dout.sizes() == q.sizes() && out.sizes() == q.sizes(), "flash_attn_bwd: dout / out must match q");
  d9d::FlashAttnArgs a = fa_args(q, k, v, scale, window_left, window_right, softcap, c10::nullopt, cu_q, cu_k, max_q, max_k);
  Tensor delta = fa_stats_like(a, q);
  TORCH_CHECK(lse.numel() == delta.numel(), "flash_attn_bwd: lse has the wrong shape");
  Tensor dq = at::empty_like(q), dk = at::empty_like(k), dv = at::empty_like(v);
  a.out = out.data_ptr();
  a.dout = dout.data_ptr();
  a.lse = lse.data_ptr<float>();
  a.delta = delta.data_ptr<float>();
  a.dq = dq.data_ptr(); a.dk = dk.data_ptr(); a.dv = dv.data_ptr();
  d9d::flash_attn_bwd_delta(a, cur_stream());
  if (dlse.has_value()) delta.sub_(dlse->reshape(delta.sizes()));  // d(lse)/dS = P folds into the row statistic
  d9d::flash_attn_bwd(a, cur_stream());
  return {dq, dk, dv, delta};
}

// ------------------------------------------------------------------ fused q/k RMSNorm + RoPE
static void check_qk(const Tensor& q, const Tensor& k) {
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && q.dim() == 3 && k.dim() == 3 && q.scalar_type() == at::kBFloat16 &&
              k.scalar_type() == at::kBFloat16, "qk_norm_rope: q/k must be bf16 [T, H, D]");
  TORCH_CHECK(q.stride(2) == 1 && k.stride(2) == 1 && q.stride(1) == q.size(2) && k.stride(1) == k.size(2) &&
              q.size(0) == k.size(0) && q.size(2) == k.size(2), "qk_norm_rope: heads must be densely packed");
}

std::tuple<Tensor, Tensor, Tensor> qk_norm_rope_fwd(const Tensor& q, const Tensor& k, const Tensor& wq, const Tensor& wk,
                                                    const Tensor& cos_t, const Tensor& sin_t, double eps, bool zero_centered,
                                                    int64_t style) {
  check_qk(q, k);
  CHECK_CUDA_CONTIG(wq); CHECK_CUDA_CONTIG(wk); CHECK_CUDA_CONTIG(cos_t); CHECK_CUDA_CONTIG(sin_t);
  TORCH_CHECK(wq.scalar_type() == at::kBFloat16 && wk.scalar_type() == at::kBFloat16 && cos_t.scalar_type() == at::kFloat &&
              sin_t.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t T = q.size(0), Hq = q.size(1), Hk = k.size(1), D = q.size(2), rope_dim = cos_t.size(-1);
  TORCH_CHECK(cos_t.numel() == T * rope_dim && sin_t.numel() == T * rope_dim && wq.numel() == D && wk.numel() == D);
  Tensor qo = at::empty({T, Hq, D}, q.options()), ko = at::empty({T, Hk, D}, k.options());
  Tensor inv = at::empty({T, Hq + Hk}, q.options().dtype(at::kFloat));
  d9d::qk_norm_rope_fwd(q.data_ptr(), k.data_ptr(), wq.data_ptr(), wk.data_ptr(), cos_t.data_ptr<float>(), sin_t.data_ptr<float>(), T,
                        static_cast<int>(Hq), static_cast<int>(Hk), static_cast<int>(D), static_cast<int>(rope_dim), q.stride(0),
                        k.stride(0), static_cast<float>(eps), zero_centered, static_cast<int>(style), qo.data_ptr(), ko.data_ptr(),
                        inv.data_ptr<float>(), cur_stream());
  return {qo, ko, inv};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> qk_norm_rope_bwd(const Tensor& dq_out, const Tensor& dk_out, const Tensor& q, const Tensor& k,
                                                            const Tensor& wq, const Tensor& wk, const Tensor& cos_t,
                                                            const Tensor& sin_t, const Tensor& inv_rms, bool zero_centered,
                                                            int64_t style) {
  check_qk(q, k);
  CHECK_CUDA_CONTIG(dq_out); CHECK_CUDA_CONTIG(dk_out); CHECK_CUDA_CONTIG(inv_rms);
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t T = q.size(0), Hq = q.size(1), Hk = k.size(1), D = q.size(2), rope_dim = cos_t.size(-1);
  Tensor dq = at::empty({T, Hq, D}, q.options()), dk = at::empty({T, Hk, D}, k.options());
  const int grid = d9d::qk_norm_rope_grid(T, static_cast<int>(Hq + Hk));
  auto fopt = q.options().dtype(at::kFloat);
  Tensor partial = at::empty({grid, 2, D}, fopt), dwq = at::empty({D}, fopt), dwk = at::empty({D}, fopt);
  d9d::qk_norm_rope_bwd(dq_out.data_ptr(), dk_out.data_ptr(), q.data_ptr(), k.data_ptr(), wq.data_ptr(), wk.data_ptr(),
                        cos_t.data_ptr<float>(), sin_t.data_ptr<float>(), inv_rms.data_ptr<float>(), T, static_cast<int>(Hq),
                        static_cast<int>(Hk), static_cast<int>(D), static_cast<int>(rope_dim), q.stride(0), k.stride(0), zero_centered,
                        static_cast<int>(style), dq.data_ptr(), dk.data_ptr(), partial.data_ptr<float>(), dwq.data_ptr<float>(),
                        dwk.data_ptr<float>(), cur_stream());
  return {dq, dk, dwq, dwk};
}

// ------------------------------------------------------------------ MoE router
std::tuple<Tensor, Tensor> router_topk_fwd(const Tensor& logits, const c10::optional<Tensor>& bias, int64_t k, bool renorm) {
  CHECK_CUDA_CONTIG(logits);
  TORCH_CHECK(logits.scalar_type() == at::kBFloat16 && logits.dim() == 2);
  c10::cuda::CUDAGuard guard(logits.device());
  const int64_t T = logits.size(0), E = logits.size(1);
  const float* b = nullptr;
  if (bias.has_value()) {
    TORCH_CHECK(bias->is_cuda() && bias->is_contiguous() && bias->scalar_type() == at::kFloat && bias->numel() == E);
    b = bias->data_ptr<float>();
  }
  Tensor idx = at::empty({T, k}, logits.options().dtype(at::kLong));
  Tensor probs = at::empty({T, k}, logits.options().dtype(at::kFloat));
  d9d::router_topk_fwd(logits.data_ptr(), b, T, static_cast<int>(E), static_cast<int>(k), renorm,
                       reinterpret_cast<long long*>(idx.data_ptr<int64_t>()), probs.data_ptr<float>(), cur_stream());
  return {idx, probs};
}

Tensor router_topk_bwd(const Tensor& logits, const Tensor& idx, const Tensor& dprobs, bool renorm) {
  CHECK_CUDA_CONTIG(logits); CHECK_CUDA_CONTIG(idx); CHECK_CUDA_CONTIG(dprobs);
  TORCH_CHECK(logits.scalar_type() == at::kBFloat16 && idx.scalar_type() == at::kLong && dprobs.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(logits.device());
  const int64_t T = logits.size(0), E = logits.size(1), k = idx.size(1);
  Tensor out = at::empty_like(logits);
  d9d::router_topk_bwd(logits.data_ptr(), reinterpret_cast<const long long*>(idx.data_ptr<int64_t>()), dprobs.data_ptr<float>(), T,
                       static_cast<int>(E), static_cast<int>(k), renorm, out.data_ptr(), cur_stream());
  return out;
}

// ------------------------------------------------------------------ tensor-parallel GEMMs with fused communication
// peer pointer lists come from torch.distributed._symmetric_memory (handle.buffer_ptrs) plus a byte offset.
static std::vector<const void*> to_ptrs(at::IntArrayRef v) {
  std::vector<const void*> out;
  out.reserve(v.size());
  for (int64_t p : v) out.push_back(reinterpret_cast<const void*>(p));
  return out;
}

// d[M,N] = all_gather(a_shards)[M,K] · B ; every peer holds a_shard[rows_local, K] (bf16, leading dim lda)
void gemm_ag_a(at::IntArrayRef a_peer_ptrs, int64_t rows_local, int64_t lda, int64_t block_rows, const Tensor& b, Tensor d,
               bool b_mn) {
  TORCH_CHECK(b.is_cuda() && d.is_cuda() && b.dim() == 2 && d.dim() == 2 && b.stride(1) == 1 && d.stride(1) == 1);
  TORCH_CHECK(b.scalar_type() == at::kBFloat16 && d.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard guard(d.device());
  auto ptrs = to_ptrs(a_peer_ptrs);
  d9d::GemmArgs g;
  g.mode = 0; g.a_mn = false; g.b_mn = b_mn; g.epi = 0;
  g.M = static_cast<int>(d.size(0)); g.N = static_cast<int>(d.size(1));
  g.K = static_cast<int>(b_mn ? b.size(0) : b.size(1));
  TORCH_CHECK((b_mn ? b.size(1) : b.size(0)) == g.N, "gemm_ag_a: N mismatch");
  TORCH_CHECK(rows_local * static_cast<int64_t>(ptrs.size()) == g.M, "gemm_ag_a: shards do not add up to M");
  g.B = b.data_ptr(); g.D = d.data_ptr(); g.ldb = b.stride(0); g.ldd = d.stride(0);
  g.comm = 1; g.comm_world = static_cast<int>(ptrs.size()); g.comm_block_rows = static_cast<int>(block_rows);
  g.comm_peer_ptrs = ptrs.data(); g.comm_rows_local = rows_local; g.comm_ld = lda;
  d9d::gemm_comm(g, cur_stream());
}

// d[M,N] = a[M,K] · B where a is a local buffer whose shards (owner(row) per the block mapping) arrive asynchronously:
// the GEMM starts with this rank's shard and waits on flags[owner] (int32, non-zero = landed) before touching another one
void gemm_wait_a(const Tensor& a, const Tensor& flags, int64_t rank, int64_t block_rows, const Tensor& b, Tensor d, bool b_mn,
                 int64_t spare_sms) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && d.is_cuda() && a.dim() == 2 && b.dim() == 2 && d.dim() == 2);
  TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1 && d.stride(1) == 1);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && d.scalar_type() == at::kBFloat16);
  TORCH_CHECK(flags.is_cuda() && flags.scalar_type() == at::kInt && flags.is_contiguous());
  c10::cuda::CUDAGuard guard(d.device());
  d9d::GemmArgs g;
  g.mode = 0; g.a_mn = false; g.b_mn = b_mn; g.epi = 0;
  g.M = static_cast<int>(a.size(0)); g.K = static_cast<int>(a.size(1));
  g.N = static_cast<int>(b_mn ? b.size(1) : b.size(0));
  TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == g.K && d.size(0) == g.M && d.size(1) == g.N, "gemm_wait_a: shape mismatch");
  g.A = a.data_ptr(); g.B = b.data_ptr(); g.D = d.data_ptr();
  g.lda = a.stride(0); g.ldb = b.stride(0); g.ldd = d.stride(0);
  g.comm = 5; g.comm_world = static_cast<int>(flags.numel()); g.comm_block_rows = static_cast<int>(block_rows);
  g.comm_rank = static_cast<int>(rank); g.comm_flags = flags.data_ptr<int>();
  g.comm_spare_sms = static_cast<int>(spare_sms);
  d9d::gemm_comm(g, cur_stream());
}

// d_shards[owner][rows_local, N] += (a[M,K] · B) rows owned by `owner` (bf16 reduce-add over NVLink)
void gemm_rs_d(const Tensor& a, const Tensor& b, at::IntArrayRef d_peer_ptrs, int64_t rows_local, int64_t ldd, int64_t block_rows,
               bool b_mn) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard guard(a.device());
  auto ptrs = to_ptrs(d_peer_ptrs);
  d9d::GemmArgs g;
  g.mode = 0; g.a_mn = false; g.b_mn = b_mn; g.epi = 3;
  g.M = static_cast<int>(a.size(0)); g.K = static_cast<int>(a.size(1));
  g.N = static_cast<int>(b_mn ? b.size(1) : b.size(0));
  TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == g.K, "gemm_rs_d: K mismatch");
  TORCH_CHECK(rows_local * static_cast<int64_t>(ptrs.size()) == g.M, "gemm_rs_d: shards do not add up to M");
  g.A = a.data_ptr(); g.B = b.data_ptr(); g.lda = a.stride(0); g.ldb = b.stride(0);
  g.comm = 2; g.comm_world = static_cast<int>(ptrs.size()); g.comm_block_rows = static_cast<int>(block_rows);
  g.comm_peer_ptrs = ptrs.data(); g.comm_rows_local = rows_local; g.comm_ld = ldd;
  d9d::gemm_comm(g, cur_stream());
}

// d[M,N] (+)= A^T · B over the token dim where one operand ([tokens, M] or [tokens, N], MN-major) is sharded by tokens
void gemm_ag_k(const Tensor& local, at::IntArrayRef peer_ptrs, bool peer_is_a, int64_t rows_local, int64_t peer_ld,
               int64_t block_rows, Tensor d, bool accumulate) {
  TORCH_CHECK(local.is_cuda() && d.is_cuda() && local.dim() == 2 && d.dim() == 2 && local.stride(1) == 1 && d.stride(1) == 1);
  TORCH_CHECK(local.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard guard(d.device());
  auto ptrs = to_ptrs(peer_ptrs);
  d9d::GemmArgs g;
  g.mode = 0; g.a_mn = true; g.b_mn = true;
  g.M = static_cast<int>(d.size(0)); g.N = static_cast<int>(d.size(1)); g.K = static_cast<int>(local.size(0));
  TORCH_CHECK(rows_local * static_cast<int64_t>(ptrs.size()) == g.K, "gemm_ag_k: shards do not add up to the token dim");
  TORCH_CHECK(local.size(1) == (peer_is_a ? g.N : g.M), "gemm_ag_k: local operand width mismatch");
  if (peer_is_a) { g.B = local.data_ptr(); g.ldb = local.stride(0); }
  else { g.A = local.data_ptr(); g.lda = local.stride(0); }
  g.D = d.data_ptr(); g.ldd = d.stride(0);
  g.epi = epi_for(d, accumulate);
  g.comm = peer_is_a ? 3 : 4; g.comm_world = static_cast<int>(ptrs.size()); g.comm_block_rows = static_cast<int>(block_rows);
  g.comm_peer_ptrs = ptrs.data(); g.comm_rows_local = rows_local; g.comm_ld = peer_ld;
  d9d::gemm_comm(g, cur_stream());
}

// ------------------------------------------------------------------ expert-parallel exchange over NVLink
void ep_push(const Tensor& x, const c10::optional<Tensor>& probs, const Tensor& dest_rank, const Tensor& dest_row, int64_t peer_ptrs_dev,
             int64_t off_x, int64_t off_p, int64_t k) {
  CHECK_CUDA_CONTIG(x); CHECK_CUDA_CONTIG(dest_rank); CHECK_CUDA_CONTIG(dest_row);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && x.dim() == 2 && dest_rank.scalar_type() == at::kInt && dest_row.scalar_type() == at::kInt);
  TORCH_CHECK(dest_rank.numel() == x.size(0) * k && dest_row.numel() == x.size(0) * k, "ep_push: one destination per (token, slot) pair");
  c10::cuda::CUDAGuard guard(x.device());
  const float* pr = nullptr;
  if (probs.has_value()) {
    TORCH_CHECK(probs->is_cuda() && probs->is_contiguous() && probs->scalar_type() == at::kFloat && probs->numel() == dest_row.numel());
    pr = probs->data_ptr<float>();
  }
  d9d::ep_push(x.data_ptr(), pr, dest_rank.data_ptr<int>(), dest_row.data_ptr<int>(), reinterpret_cast<void* const*>(peer_ptrs_dev), off_x,
               off_p, dest_row.numel(), static_cast<int>(k), static_cast<int>(x.size(1)), cur_stream());
}

std::tuple<Tensor, Tensor> ep_pull_sum(int64_t peer_ptrs_dev, int64_t off_y, int64_t off_dp, const Tensor& dest_rank, const Tensor& dest_row,
                                       int64_t T, int64_t k, int64_t H, bool with_dprobs) {
  CHECK_CUDA_CONTIG(dest_rank); CHECK_CUDA_CONTIG(dest_row);
  TORCH_CHECK(dest_rank.scalar_type() == at::kInt && dest_row.scalar_type() == at::kInt && dest_row.numel() == T * k);
  c10::cuda::CUDAGuard guard(dest_row.device());
  Tensor y = at::empty({T, H}, dest_row.options().dtype(at::kBFloat16));
  Tensor dprobs = at::empty({with_dprobs ? T : 0, k}, dest_row.options().dtype(at::kFloat));
  d9d::ep_pull_sum(reinterpret_cast<void* const*>(peer_ptrs_dev), off_y, off_dp, dest_rank.data_ptr<int>(), dest_row.data_ptr<int>(),
                   y.data_ptr(), with_dprobs ? dprobs.data_ptr<float>() : nullptr, T, static_cast<int>(k), static_cast<int>(H), cur_stream());
  return {y, dprobs};
}

// ------------------------------------------------------------------ NVLink data-parallel optimizer -----------
// `peer_ptrs_dev` / `multicast_ptr` come from torch.distributed._symmetric_memory (buffer_ptrs_dev, multicast_ptr).
void nvl_reduce_shard_(Tensor own_grad, int64_t peer_ptrs_dev, int64_t multicast_ptr, int64_t begin, int64_t end,
                       int64_t world, int64_t rank, Tensor sumsq) {
  CHECK_CUDA_CONTIG(own_grad); CHECK_CUDA_CONTIG(sumsq);
  TORCH_CHECK(own_grad.scalar_type() == at::kFloat && sumsq.scalar_type() == at::kFloat);
  TORCH_CHECK(0 <= begin && begin <= end && end <= own_grad.numel(), "nvl_reduce_shard_: bad shard bounds");
  c10::cuda::CUDAGuard guard(own_grad.device());
  d9d::nvl_reduce_shard(reinterpret_cast<const float* const*>(peer_ptrs_dev), reinterpret_cast<const float*>(multicast_ptr),
                        own_grad.data_ptr<float>(), begin, end, static_cast<int>(world), static_cast<int>(rank),
                        sumsq.data_ptr<float>(), cur_stream());
}

void nvl_adamw_shard_(Tensor own_param, const Tensor& own_grad, Tensor exp_avg, Tensor exp_avg_sq, int64_t peer_ptrs_dev,
                      int64_t multicast_ptr, int64_t begin, int64_t end, int64_t world, int64_t rank, double lr,
                      double beta1, double beta2, double eps, double weight_decay, double bias_corr1, double bias_corr2,
                      int64_t seed, const c10::optional<Tensor>& grad_scale) {
  CHECK_CUDA_CONTIG(own_param); CHECK_CUDA_CONTIG(own_grad); CHECK_CUDA_CONTIG(exp_avg); CHECK_CUDA_CONTIG(exp_avg_sq);
  TORCH_CHECK(own_param.scalar_type() == at::kBFloat16 && own_grad.scalar_type() == at::kFloat);
  TORCH_CHECK(exp_avg.scalar_type() == exp_avg_sq.scalar_type() &&
              (exp_avg.scalar_type() == at::kBFloat16 || exp_avg.scalar_type() == at::kFloat));
  TORCH_CHECK(0 <= begin && begin <= end && end <= own_param.numel() && end <= own_grad.numel(), "nvl_adamw_shard_: bad shard bounds");
  TORCH_CHECK(exp_avg.numel() >= end - begin && exp_avg_sq.numel() >= end - begin, "nvl_adamw_shard_: state smaller than shard");
  c10::cuda::CUDAGuard guard(own_param.device());
  const float* gs = nullptr;
  if (grad_scale.has_value()) {
    TORCH_CHECK(grad_scale->is_cuda() && grad_scale->scalar_type() == at::kFloat);
    gs = grad_scale->data_ptr<float>();
  }
  d9d::nvl_adamw_shard(reinterpret_cast<void* const*>(peer_ptrs_dev), reinterpret_cast<void*>(multicast_ptr), own_param.data_ptr(),
                       own_grad.data_ptr<float>(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), begin, end,
                       static_cast<int>(world), static_cast<int>(rank), static_cast<float>(lr), static_cast<float>(beta1),
                       static_cast<float>(beta2), static_cast<float>(eps), static_cast<float>(weight_decay),
                       static_cast<float>(bias_corr1), static_cast<float>(bias_corr2), static_cast<uint64_t>(seed), gs,
                       exp_avg.scalar_type() == at::kBFloat16, cur_stream());
}

TORCH_LIBRARY(d9d_b200, m) {
  m.def("gemm(Tensor a, Tensor b, Tensor(a!) d, bool a_mn, bool b_mn, bool accumulate, int ctas=0) -> ()");
  m.def("gemm_set_pair_mode(int mode) -> int", &gemm_set_pair_mode);
  m.def("gemm_grouped_m(Tensor a, Tensor b, Tensor(a!) d, Tensor tile_group, bool b_mn, bool accumulate=False) -> ()");
  m.def("gemm_grouped_k(Tensor a, Tensor b, Tensor(a!) d, Tensor group_offsets, bool accumulate) -> ()");
  m.def("ce_forward(Tensor h, Tensor w, Tensor target, int ignore_index) -> (Tensor, Tensor)");
  m.def("ce_forward_ex(Tensor h, Tensor w, Tensor target, int ignore_index, Tensor? bias=None, float softcap=0.0, int col_offset=0) -> (Tensor, Tensor, Tensor)");
  m.def("ce_dlogits(Tensor h, Tensor w, Tensor target, Tensor lse, Tensor grad, Tensor(a!) out, int ignore_index, Tensor? bias=None, float softcap=0.0, int col_offset=0) -> ()");
  m.def("quantize_rowwise(Tensor x) -> (Tensor, Tensor)");
  m.def("quantize_colwise_t(Tensor x) -> (Tensor, Tensor)");
  m.def("quantize_mx(Tensor x) -> (Tensor, Tensor)");
  m.def("gemm_fp8(Tensor a, Tensor b, Tensor? scale_a, Tensor? scale_b, float scalar, Tensor(a!) d) -> ()");
  m.def("gemm_mxfp8(Tensor a, Tensor sfa, Tensor b, Tensor sfb, Tensor(a!) d) -> ()");
  m.def("rms_norm_fwd(Tensor x, Tensor w, float eps, bool zero_centered) -> (Tensor, Tensor)");
  m.def("rms_norm_bwd(Tensor dout, Tensor x, Tensor w, Tensor inv_rms, bool zero_centered) -> (Tensor, Tensor)");
  m.def("silu_mul_fwd(Tensor x, Tensor y) -> Tensor");
  m.def("silu_mul_bwd(Tensor dout, Tensor x, Tensor y) -> (Tensor, Tensor)");
  m.def("silu_mul_probs_fwd(Tensor x, Tensor y, Tensor probs, Tensor? valid_rows=None) -> Tensor");
  m.def("silu_mul_probs_bwd(Tensor dout, Tensor x, Tensor y, Tensor probs, Tensor? valid_rows=None) -> (Tensor, Tensor, Tensor)");
  m.def("sr_copy_(Tensor(a!) dst, Tensor src, int seed) -> ()");
  m.def(
      "adamw_sr_multi_(Tensor metas, Tensor block_map, float lr, float beta1, float beta2, float eps, float "
      "weight_decay, int step, int seed, Tensor? grad_scale, bool grad_bf16, bool state_bf16) -> ()");
  m.def("rope_apply(Tensor x, Tensor cos_cache, Tensor sin_cache, Tensor pos, int style, bool inverse) -> Tensor");
  m.def("moe_build_layout(Tensor topk_ids, int num_experts, int align, int capacity) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("moe_permute(Tensor x, Tensor? probs, Tensor row_map, Tensor counts, Tensor seg_offsets, int capacity) -> (Tensor, Tensor)");
  m.def("moe_gather(Tensor yp, Tensor? dpp, Tensor row_map, int T, int k) -> (Tensor, Tensor)");
  m.def("sumsq_accumulate_(Tensor x, Tensor(a!) out) -> ()");
  m.def("moe_zero_pad(Tensor(a!) xp, Tensor? pp, Tensor counts, Tensor seg_offsets) -> ()");
  m.def("flash_attn_fwd(Tensor q, Tensor k, Tensor v, float scale, int window_left, int window_right, float softcap, Tensor? sink, Tensor? cu_q, Tensor? cu_k, int max_q, int max_k, int variant) -> (Tensor, Tensor)");
  m.def("flash_attn_bwd(Tensor dout, Tensor q, Tensor k, Tensor v, Tensor out, Tensor lse, float scale, int window_left, int window_right, float softcap, Tensor? cu_q, Tensor? cu_k, int max_q, int max_k, Tensor? dlse) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("qk_norm_rope_fwd(Tensor q, Tensor k, Tensor wq, Tensor wk, Tensor cos_t, Tensor sin_t, float eps, bool zero_centered, "
        "int style) -> (Tensor, Tensor, Tensor)");
  m.def("qk_norm_rope_bwd(Tensor dq_out, Tensor dk_out, Tensor q, Tensor k, Tensor wq, Tensor wk, Tensor cos_t, Tensor sin_t, "
        "Tensor inv_rms, bool zero_centered, int style) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("router_topk_fwd(Tensor logits, Tensor? bias, int k, bool renorm) -> (Tensor, Tensor)");
  m.def("router_topk_bwd(Tensor logits, Tensor idx, Tensor dprobs, bool renorm) -> Tensor");
  m.def("gemm_ag_a(int[] a_peer_ptrs, int rows_local, int lda, int block_rows, Tensor b, Tensor(a!) d, bool b_mn) -> ()");
  m.def("gemm_wait_a(Tensor a, Tensor flags, int rank, int block_rows, Tensor b, Tensor(a!) d, bool b_mn, int spare_sms=0) -> ()");
  m.def("gemm_rs_d(Tensor a, Tensor b, int[] d_peer_ptrs, int rows_local, int ldd, int block_rows, bool b_mn) -> ()");
  m.def("gemm_ag_k(Tensor local, int[] peer_ptrs, bool peer_is_a, int rows_local, int peer_ld, int block_rows, Tensor(a!) d, "
        "bool accumulate) -> ()");
  m.def("ep_push(Tensor x, Tensor? probs, Tensor dest_rank, Tensor dest_row, int peer_ptrs_dev, int off_x, int off_p, int k) -> ()");
  m.def("ep_pull_sum(int peer_ptrs_dev, int off_y, int off_dp, Tensor dest_rank, Tensor dest_row, int T, int k, int H, bool with_dprobs) "
        "-> (Tensor, Tensor)");
  m.def("nvl_reduce_shard_(Tensor(a!) own_grad, int peer_ptrs_dev, int multicast_ptr, int begin, int end, int world, int rank, "
        "Tensor(b!) sumsq) -> ()");
  m.def("nvl_adamw_shard_(Tensor(a!) own_param, Tensor own_grad, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, int peer_ptrs_dev, "
        "int multicast_ptr, int begin, int end, int world, int rank, float lr, float beta1, float beta2, float eps, "
        "float weight_decay, float bias_corr1, float bias_corr2, int seed, Tensor? grad_scale) -> ()");
  m.def("scale_inplace_(Tensor(a!) x, Tensor scale) -> ()");
}

TORCH_LIBRARY_IMPL(d9d_b200, CUDA, m) {
  m.impl("gemm", &gemm);
  m.impl("gemm_grouped_m", &gemm_grouped_m);
  m.impl("gemm_grouped_k", &gemm_grouped_k);
  m.impl("ce_forward", &ce_forward);
  m.impl("ce_forward_ex", &ce_forward_ex);
  m.impl("ce_dlogits", &ce_dlogits);
  m.impl("quantize_rowwise", &quantize_rowwise);
  m.impl("quantize_colwise_t", &quantize_colwise_t);
  m.impl("quantize_mx", &quantize_mx);
  m.impl("gemm_fp8", &gemm_fp8);
  m.impl("gemm_mxfp8", &gemm_mxfp8);
  m.impl("rms_norm_fwd", &rms_norm_fwd);
  m.impl("rms_norm_bwd", &rms_norm_bwd);
  m.impl("silu_mul_fwd", &silu_mul_fwd);
  m.impl("silu_mul_bwd", &silu_mul_bwd);
  m.impl("silu_mul_probs_fwd", &silu_mul_probs_fwd);
  m.impl("silu_mul_probs_bwd", &silu_mul_probs_bwd);
  m.impl("sr_copy_", &sr_copy_);
  m.impl("adamw_sr_multi_", &adamw_sr_multi_);
  m.impl("rope_apply", &rope_apply);
  m.impl("moe_build_layout", &moe_build_layout);
  m.impl("moe_permute", &moe_permute);
  m.impl("moe_gather", &moe_gather);
  m.impl("sumsq_accumulate_", &sumsq_accumulate_);
  m.impl("moe_zero_pad", &moe_zero_pad);
  m.impl("flash_attn_fwd", &flash_attn_fwd);
  m.impl("flash_attn_bwd", &flash_attn_bwd);
  m.impl("qk_norm_rope_fwd", &qk_norm_rope_fwd);
  m.impl("qk_norm_rope_bwd", &qk_norm_rope_bwd);
  m.impl("router_topk_fwd", &router_topk_fwd);
  m.impl("router_topk_bwd", &router_topk_bwd);
  m.impl("gemm_ag_a", &gemm_ag_a);
  m.impl("gemm_wait_a", &gemm_wait_a);
  m.impl("gemm_rs_d", &gemm_rs_d);
  m.impl("gemm_ag_k", &gemm_ag_k);
  m.impl("ep_push", &ep_push);
  m.impl("ep_pull_sum", &ep_pull_sum);
  m.impl("nvl_reduce_shard_", &nvl_reduce_shard_);
  m.impl("nvl_adamw_shard_", &nvl_adamw_shard_);
  m.impl("scale_inplace_", &scale_inplace_);
}
