// RMS normalisation fused with the static INT8 quantiser (the prologue in front
// of q/k/v resp. w1/w2; one read of the hidden state instead of read + write + read).
//
//   RMSN.forward (fake_quant/module_util.py:42-61):
//     x = x.to(float32); variance = x.pow(2).sum(-1, keepdim=True) / mean_dim
//     x = x * torch.rsqrt(variance + eps); return x.to(float16)
//   followed by ActQuantizer / UniformQuantizer.quant of the fp16 result (quant_utils.py:151-167,
//   quantizer/uniform.py:28-31) with the token's MSQ scale.
//
// One warp per token: the row stays in registers between the sum of squares and the scaling
// (C <= 8192), every element is read once.  Algorithmic bytes per element: 2 read + 1 written
// (codes) [+ 2 for the fp16 output].  The fp32 sum runs over four partial sums per lane and a
// butterfly across the lanes -- a different order than torch's reduction, which moves the fp16
// result by one ulp on a few 1e-4 of the elements (tests/test_gpu_kernels.py states the bar).
#include "mq_common.cuh"

namespace mq {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 1 / sqrt(variance + eps) with IEEE operations (torch's CPU rsqrt is 1 / sqrt)
__device__ __forceinline__ float rmsn_factor(float sumsq, float mean_dim, float eps) {
  return __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(sumsq, mean_dim), eps)));
}

__device__ __forceinline__ uint4 scale8_f16(const uint4& v, float r) {
  float f[8];
  unpack8(v, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __fmul_rn(f[i], r);
  return pack8_f16(f);
}

// NCH 16-byte chunks per lane (C = 256 * NCH); NARROW: division-free quantiser for both scales
template <int NCH, bool NARROW>
__device__ __forceinline__ void rmsn_rows(const __half* __restrict__ x, int64_t ldx, int8_t* __restrict__ q,
                                          __half* __restrict__ y, float* __restrict__ srow_out,
                                          const uint8_t* __restrict__ modality, const QScale& q_text, const QScale& q_vis,
                                          int64_t M, int C, int n_out, float mean_dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const QNarrow n_text = make_qnarrow(q_text), n_vis = make_qnarrow(q_vis);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps) {
    const __half* xr = x + m * ldx;
    uint4 v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = ldg_stream(xr + (c * 32 + lane) * 8);
    const bool vis = modality != nullptr && __ldg(modality + m) != 0;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float f[8];
      unpack8(v[c], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i & 3] = __fmaf_rn(f[i], f[i], acc[i & 3]);  // (x*x is exact in fp32)
    }
    const float r = rmsn_factor(warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3])), mean_dim, eps);
    const QScale& qs = vis ? q_vis : q_text;
    const QNarrow& qn = vis ? n_vis : n_text;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint4 h = scale8_f16(v[c], r);  // x * rsqrt(.) -> fp16
      const int col = (c * 32 + lane) * 8;
      if (y != nullptr) stg_stream(y + m * (int64_t)C + col, h);
      if (q != nullptr) {
        const uint2 codes = NARROW ? quant8_narrow(h, qn) : quant8_bracket(h, qs);
        *reinterpret_cast<uint2*>(q + m * (int64_t)n_out + col) = codes;
      }
    }
    if (q != nullptr) {
      for (int col = C + lane * 8; col < n_out; col += 256)  // zero padding up to the GEMM's K
        *reinterpret_cast<uint2*>(q + m * (int64_t)n_out + col) = make_uint2(0u, 0u);
      if (lane == 0 && srow_out != nullptr) srow_out[m] = qs.s;
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(128)
rmsn_quant_kernel(const __half* __restrict__ x, int64_t ldx, int8_t* __restrict__ q, __half* __restrict__ y,
                  float* __restrict__ srow_out, const uint8_t* __restrict__ modality, const float* __restrict__ scales,
                  int64_t M, int C, int n_out, float mean_dim, float eps) {
  QScale q_text = {}, q_vis = {};
  bool narrow = true;
  if (q != nullptr) {
    q_text = make_qscale(__ldg(scales));
    q_vis = make_qscale(__ldg(scales + 1));
    narrow = q_text.narrow && q_vis.narrow;
  }
  if (narrow)
    rmsn_rows<NCH, true>(x, ldx, q, y, srow_out, modality, q_text, q_vis, M, C, n_out, mean_dim, eps);
  else
    rmsn_rows<NCH, false>(x, ldx, q, y, srow_out, modality, q_text, q_vis, M, C, n_out, mean_dim, eps);
}

// any C % 8 == 0: two passes over the row (the second one hits L1 / L2)
__global__ void __launch_bounds__(128)
rmsn_quant_generic_kernel(const __half* __restrict__ x, int64_t ldx, int8_t* __restrict__ q, __half* __restrict__ y,
                          float* __restrict__ srow_out, const uint8_t* __restrict__ modality,
                          const float* __restrict__ scales, int64_t M, int C, int n_out, float mean_dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  QScale q_text = {}, q_vis = {};
  if (q != nullptr) {
    q_text = make_qscale(__ldg(scales));
    q_vis = make_qscale(__ldg(scales + 1));
  }
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps) {
    const __half* xr = x + m * ldx;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int col = lane * 8; col < C; col += 256) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(xr + col), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i & 3] = __fmaf_rn(f[i], f[i], acc[i & 3]);
    }
    const float r = rmsn_factor(warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3])), mean_dim, eps);
    const bool vis = modality != nullptr && __ldg(modality + m) != 0;
    const QScale& qs = vis ? q_vis : q_text;
    for (int col = lane * 8; col < n_out; col += 256) {
      if (col < C) {
        const uint4 h = scale8_f16(*reinterpret_cast<const uint4*>(xr + col), r);
        if (y != nullptr) stg_stream(y + m * (int64_t)C + col, h);
        if (q != nullptr) *reinterpret_cast<uint2*>(q + m * (int64_t)n_out + col) = quant8(h, qs);
      } else if (q != nullptr) {
        *reinterpret_cast<uint2*>(q + m * (int64_t)n_out + col) = make_uint2(0u, 0u);
      }
    }
    if (q != nullptr && lane == 0 && srow_out != nullptr) srow_out[m] = qs.s;
  }
}

}  // namespace mq

extern "C" int mq_rmsn_quant_i8(const void* x, int64_t ldx, int8_t* q, void* y_f16, float* srow_out,
                                const uint8_t* modality, const float* scales, int64_t M, int C, int n_out,
                                float mean_dim, float eps, void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;
  MQ_REQUIRE(x != nullptr && (q != nullptr || y_f16 != nullptr), "mq_rmsn_quant_i8: null pointer");
  MQ_REQUIRE(q == nullptr || scales != nullptr, "mq_rmsn_quant_i8: codes requested without scales");
  MQ_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C, "mq_rmsn_quant_i8: bad shape (C=%d ldx=%lld)", C,
             (long long)ldx);
  MQ_REQUIRE(q == nullptr || (n_out >= C && n_out % 8 == 0), "mq_rmsn_quant_i8: n_out=%d must be >= C and a multiple of 8", n_out);
  MQ_REQUIRE(mean_dim > 0.f, "mq_rmsn_quant_i8: mean_dim must be positive");
  MQ_REQUIRE(aligned16(x) && aligned16(y_f16) && (reinterpret_cast<uintptr_t>(q) & 7u) == 0,
             "mq_rmsn_quant_i8: pointers must be 16-byte (codes: 8-byte) aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int wpb = 4;
  int64_t blocks = (M + wpb - 1) / wpb;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const __half* xp = (const __half*)x;
  __half* yp = (__half*)y_f16;
#define MQ_RMSN_CASE(NCH)                                                                                          \
  case NCH:                                                                                                        \
    rmsn_quant_kernel<NCH><<<(int)blocks, wpb * 32, 0, st>>>(xp, ldx, q, yp, srow_out, modality, scales, M, C, n_out, \
                                                             mean_dim, eps);                                       \
    break;
  const int nch = (C % 256 == 0) ? C / 256 : 0;
  switch (nch) {
    MQ_RMSN_CASE(4)
    MQ_RMSN_CASE(8)
    MQ_RMSN_CASE(14)
    MQ_RMSN_CASE(16)
    MQ_RMSN_CASE(20)
    MQ_RMSN_CASE(32)
    default:
      rmsn_quant_generic_kernel<<<(int)blocks, wpb * 32, 0, st>>>(xp, ldx, q, yp, srow_out, modality, scales, M, C, n_out,
                                                                   mean_dim, eps);
  }
#undef MQ_RMSN_CASE
  MQ_LAUNCH_CHECK("mq_rmsn_quant_i8");
  return MQ_OK;
}
