// fp8 (e4m3) GEMM launchers + the quantisation kernels that produce their operands.
//   quantize_rowwise     x[M,K] bf16 -> q[M,K] e4m3, scale[M]            (scale = amax / 448 per row)
//   quantize_colwise_t   x[R,C] bf16 -> qT[C,R] e4m3, scale[C]           (transposed copy, scale per column of x)
//   quantize_mx          x[M,K] bf16 -> q[M,K] e4m3, UE8M0 scale blocks  (one scale per 32 K-elements, tcgen05 block layout)
//   gemm_fp8 / gemm_mxfp8  D[M,N] bf16 = scaled A·B^T on kind::f8f6f4 / kind::mxf8f6f4.block_scale
#include <cuda_fp8.h>

#include "gemm_fp8_sm100.cuh"
#include "gemm_host.cuh"

namespace d9d {
namespace {

constexpr float E4M3_MAX = 448.f;

__device__ __forceinline__ uint16_t cvt_e4m3x2(float lo, float hi) {
  return static_cast<uint16_t>(__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3));
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

__device__ __forceinline__ uint2 quant8(const float (&v)[8], float inv) {
  uint2 o;
  o.x = cvt_e4m3x2(v[0] * inv, v[1] * inv) | (static_cast<uint32_t>(cvt_e4m3x2(v[2] * inv, v[3] * inv)) << 16);
  o.y = cvt_e4m3x2(v[4] * inv, v[5] * inv) | (static_cast<uint32_t>(cvt_e4m3x2(v[6] * inv, v[7] * inv)) << 16);
  return o;
}

// ---- row-wise: one warp per row, K % 8 == 0 --------------------------------------------------------------------
__global__ void quant_rowwise_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, float* __restrict__ scale,
                                     long long M, int K, long long ldx) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* xr = x + row * ldx;
  float amax = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float v[8];
    load8_bf16(xr + k, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[i]));
  }
  amax = warp_max(amax);
  const float s = amax > 0.f ? amax / E4M3_MAX : 1.f;
  const float inv = 1.f / s;
  if (lane == 0) scale[row] = s;
  uint8_t* qr = q + row * K;
  for (int k = lane * 8; k < K; k += 256) {
    float v[8];
    load8_bf16(xr + k, v);
    *reinterpret_cast<uint2*>(qr + k) = quant8(v, inv);
  }
}

// ---- column-wise + transpose -----------------------------------------------------------------------------------
// pass 1: amax per column (non-negative floats order like their bit patterns: atomicMax on the int view)
__global__ void col_amax_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ amax, long long R, int C, long long ldx,
                                int rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  float m = 0.f;
  for (long long r = r0; r < r1; ++r) m = fmaxf(m, fabsf(__bfloat162float(x[r * ldx + c])));
  atomicMax(reinterpret_cast<int*>(amax) + c, __float_as_int(m));
}

// pass 2: 64 x 64 tile through shared memory; amax[c] is turned into the scale in place by the blocks of the first tile row
__global__ void quant_colwise_t_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ amax, uint8_t* __restrict__ qt,
                                       float* __restrict__ scale, long long R, int C, long long ldx, long long ldq) {
  __shared__ uint8_t tile[64][64 + 4];
  const int c0 = blockIdx.x * 64;
  const long long r0 = static_cast<long long>(blockIdx.y) * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 256 threads: 64 columns x 4 row phases
  const int c = c0 + tx;
  float inv = 0.f;
  if (c < C) {
    const float a = amax[c];
    const float s = a > 0.f ? a / E4M3_MAX : 1.f;
    inv = 1.f / s;
    if (blockIdx.y == 0 && ty == 0) scale[c] = s;
  }
  for (int i = ty; i < 64; i += 4) {
    const long long r = r0 + i;
    float v = 0.f;
    if (r < R && c < C) v = __bfloat162float(x[r * ldx + c]) * inv;
    tile[tx][i] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
  }
  __syncthreads();
  // write qT[c0 + j, r0 .. r0+63]: 64 bytes per output row, 4 rows per pass
  for (int j = ty; j < 64; j += 4) {
    const long long r = r0 + tx;
    if (c0 + j < C && r < R) qt[static_cast<long long>(c0 + j) * ldq + r] = tile[j][tx];
  }
}

// ---- MX: one thread per 32-element group -----------------------------------------------------------------------
__global__ void quant_mx_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf, long long M,
                                int K, long long ldx) {
  const int groups_per_row = K / 32;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= M * groups_per_row) return;
  const long long row = t / groups_per_row;
  const int g = static_cast<int>(t - row * groups_per_row);
  const __nv_bfloat16* src = x + row * ldx + g * 32;
  float v[4][8];
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    load8_bf16(src + j * 8, v[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[j][i]));
  }
  // smallest power of two 2^e with amax / 2^e <= 448 = 1.75 * 2^8: with amax = m * 2^x (1 <= m < 2) that is
  // e = x - 8 (+1 if m > 1.75) - exact integer arithmetic on the float's bit pattern
  int e = -127;
  if (amax > 0.f) {
    const uint32_t bits = __float_as_uint(amax);
    e = static_cast<int>((bits >> 23) & 0xFF) - 127 - 8 + ((bits & 0x7FFFFF) > 0x600000 ? 1 : 0);
    e = max(-126, min(126, e));
  }
  const float inv = (e == -127) ? 0.f : __uint_as_float(static_cast<uint32_t>(127 - e) << 23);
  uint8_t* dst = q + row * K + g * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<uint2*>(dst + j * 8) = quant8(v[j], inv);
  const long long blk = (row / 128) * (K / 128) + (g / 4);
  const int r = static_cast<int>(row % 128);
  sf[blk * 512 + (r % 32) * 16 + (r / 32) * 4 + (g % 4)] = static_cast<uint8_t>(e + 127);
}

CUtensorMap make_tmap_u8_2d(const void* base, uint64_t d0, uint64_t d1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1) {
  CUtensorMap m;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (strides[0] & 15) != 0)
    throw std::runtime_error("d9d gemm_fp8: operand base / row pitch must be 16-byte aligned for TMA");
  CUresult r = gemm::get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("d9d gemm_fp8: cuTensorMapEncodeTiled failed: " + std::to_string(r));
  return m;
}

template <int BLOCK_N, int SCALING>
void launch_fp8(const Fp8GemmArgs& a, cudaStream_t stream) {
  using C = fp8::Cfg<BLOCK_N, SCALING>;
  auto kern = fp8::gemm_fp8_kernel<BLOCK_N, SCALING>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    configured = true;
  }
  CUtensorMap ta = make_tmap_u8_2d(a.A, a.K, a.M, a.lda, fp8::BLOCK_K, fp8::BLOCK_M);
  CUtensorMap tb = make_tmap_u8_2d(a.B, a.K, a.N, a.ldb, fp8::BLOCK_K, BLOCK_N);
  CUtensorMap td;
  if (!gemm::make_tmap_out(&td, a.D, true, a.N, a.M, 1, a.ldd, 0))
    throw std::runtime_error("d9d gemm_fp8: output must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
  fp8::Params p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.scale_a = a.scale_a; p.scale_b = a.scale_b; p.scale_scalar = a.scale_scalar;
  p.sfa = a.sfa; p.sfb = a.sfb;
  const long long tiles = static_cast<long long>((a.M + fp8::BLOCK_M - 1) / fp8::BLOCK_M) * ((a.N + BLOCK_N - 1) / BLOCK_N);
  if (tiles <= 0) return;
  const int grid = static_cast<int>(tiles < gemm::sm_count() ? tiles : gemm::sm_count());
  kern<<<grid, fp8::NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, td, p);
}

}  // namespace

void quantize_rowwise_e4m3(const void* x, void* q, float* scale, long long M, int K, long long ldx, cudaStream_t stream) {
  if (M <= 0) return;
  if (K % 8 != 0) throw std::runtime_error("d9d quantize_rowwise: K must be a multiple of 8");
  const int warps = 4;
  quant_rowwise_kernel<<<static_cast<unsigned>((M + warps - 1) / warps), warps * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(q), scale, M, K, ldx);
}

void quantize_colwise_t_e4m3(const void* x, void* qt, float* scale, float* amax_scratch, long long R, int C, long long ldx,
                             long long ldq, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return;
  cudaMemsetAsync(amax_scratch, 0, sizeof(float) * C, stream);
  const int rows_per_block = 256;
  dim3 g1((C + 127) / 128, static_cast<unsigned>((R + rows_per_block - 1) / rows_per_block));
  col_amax_kernel<<<g1, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), amax_scratch, R, C, ldx, rows_per_block);
  dim3 g2((C + 63) / 64, static_cast<unsigned>((R + 63) / 64));
  quant_colwise_t_kernel<<<g2, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), amax_scratch, static_cast<uint8_t*>(qt),
                                                 scale, R, C, ldx, ldq);
}

void quantize_mx_e4m3(const void* x, void* q, void* sf, long long M, int K, long long ldx, cudaStream_t stream) {
  if (M <= 0) return;
  if (K % 128 != 0) throw std::runtime_error("d9d quantize_mx: K must be a multiple of 128");
  const long long n = M * (K / 32);
  quant_mx_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                             static_cast<uint8_t*>(q), static_cast<uint8_t*>(sf), M, K, ldx);
}

void gemm_fp8(const Fp8GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return;
  if (a.block_scaled) {
    if (a.K % 128 != 0 || a.sfa == nullptr || a.sfb == nullptr)
      throw std::runtime_error("d9d gemm_mxfp8: K must be a multiple of 128 and both scale-factor tensors are required");
    if (a.N <= 128) launch_fp8<128, fp8::MX>(a, stream);
    else launch_fp8<256, fp8::MX>(a, stream);
  } else {
    if (a.N <= 128) launch_fp8<128, fp8::ROWCOL>(a, stream);
    else launch_fp8<256, fp8::ROWCOL>(a, stream);
  }
}

}  // namespace d9d
