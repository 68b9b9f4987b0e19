// W4A8 GEMM for sm_100a -- C-ABI entry points and kernel dispatch.
//
//   out[m,n] = fp16( fp32(sum_k qa[m,k] * w[n,k]) * (s_a[m]*s_w[n]) (+ x0[m]*w0[n]) (+ bias[n]) )
//
// Replaces ActQuantWrapper's `self.module(x)` (fake_quant/quant_utils.py:384) and the fp32
// L1/L2 split (:374-376).  The integer sum is exact; the epilogue is specified operation by
// operation (oracle/mq_oracle.py:qlinear_int) and reproduced bit for bit by every kernel:
//
//   mq_gemm_w4a8_v4.cu   256 x 256 tile, one weight conversion shared by two M halves (large M)
//   mq_gemm_w4a8_v5.cu   128 channels x 256 tokens, converted weights as the TMEM-resident A operand
//                        (mid-sized M: one or two waves of tiles)
//   mq_gemm_w4a8_v2.cu   128 x 256 tile, weights converted into shared memory (forced / split-K base)
//   mq_gemm_w4a8_sk.cu   split-K flavour of the 128 x 256 kernel (mid-sized M)
//   mq_gemv_w4a8.cu      weight-streaming dp4a kernel (decode)
//
// INT8 activations x INT4 weights on tcgen05 (kind::i8), accumulators in TMEM, activations by
// TMA, packed INT4 weights by one bulk copy per k-block, unpacked to INT8 into the UMMA
// shared-memory layout by converter warps.  (The first-generation kernels live in
// tools/experiments/superseded/ and are not part of the library.)
#include <cuda.h>

#include <atomic>

#include "mq_gemm_common.cuh"

namespace mq {
namespace gemm {
using namespace mq::tc;

constexpr int BM = 128;  // rows of the activation TMA box
constexpr int BN = 256;  // output channels per packed-weight tile
constexpr int BK = 128;  // int8 elements per k-block

// test / tooling override of the kernel choice (mq_debug_force_gemm_kernel); 0 = automatic
static std::atomic<int> g_force_kernel{0};

// Which kernel the dispatcher runs for a shape (no environment overrides):
//   1 weight-streaming dp4a kernel (decode)      2 split-K tcgen05 kernel (few tiles, long K)
//   4 256 x 256 tcgen05 kernel (large M)         5 128-channel x 256-token kernel, weights in TMEM
//   (3 = the 128 x 256 kernel: superseded by 5 for every shape measured, reachable by forcing only)
static int plan_kernel(int64_t M, int64_t N, int64_t K, size_t workspace_bytes) {
  const int64_t n_tiles = (N + BN - 1) / BN, k_tiles = K / BK;
  const bool gemv_shape = M <= 8 || (M <= 16 && n_tiles < 32);
  if (gemv_shape && workspace_bytes > 0 && workspace_bytes >= gemv_workspace_bytes(N)) return 1;
  // Wave model fitted to tools/latency.py (hot code, cold weights; us, K in k-blocks;
  // profiles/r2m_gemm_latency.log).  A persistent CTA overlaps the loads of its next tile with the
  // epilogue of the current one, so the first wave costs more than each further one:
  //   256 x 256 tiles                     first 0.60 K + 5.5    further 0.60 K + 3.3
  //   128 x 256 tiles (weights in TMEM)   first 0.28 K + 11.4   further 0.28 K + 5.6
  //   a split-K item                      0.44 K/split + 17 (9 of it the reduction)
  // (one 1024-token request: q|k|v = 2 waves of the first kernel, 47 us, against 3 of the second, 50 us)
  const int64_t sms = num_sms();
  const int64_t waves4 = (((M + 255) / 256) * n_tiles + sms - 1) / sms;
  const int64_t waves5 = (((M + 255) / 256) * ((N + 127) / 128) + sms - 1) / sms;
  const double cost4 = 0.60 * k_tiles + 5.5 + (double)(waves4 - 1) * (0.60 * k_tiles + 3.3);
  const double cost5 = 0.28 * k_tiles + 11.4 + (double)(waves5 - 1) * (0.28 * k_tiles + 5.6);
  int best = cost4 < cost5 ? 4 : 5;
  const double cost = cost4 < cost5 ? cost4 : cost5;
  const int split = splitk_factor(M, N, K);
  if (workspace_bytes > 0 && split >= 2 && workspace_bytes >= splitk_workspace_bytes(M, N, K)) {
    const double cost_sk = 0.44 * (double)((k_tiles + split - 1) / split) + 17.0;
    if (cost_sk < cost) best = 2;
  }
  return best;
}
}  // namespace gemm
}  // namespace mq

extern "C" int mq_w4a8_gemm_plan(int64_t M, int64_t N, int64_t K, size_t workspace_bytes) {
  if (M <= 0 || N <= 0 || K <= 0 || K % mq::gemm::BK != 0) return 0;
  return mq::gemm::plan_kernel(M, N, K, workspace_bytes);
}

extern "C" size_t mq_w4a8_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  // one call uses at most one of the two (both leave the workspace zeroed), so they may alias
  const size_t a = M <= 16 ? mq::gemv_workspace_bytes(N) : 0;
  const size_t b = K % mq::gemm::BK == 0 ? mq::gemm::splitk_workspace_bytes(M, N, K) : 0;
  return a > b ? a : b;
}

extern "C" int mq_w4a8_gemm(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                            const float* s_w, const void* bias, const void* x0, const void* w0, void* out,
                            int64_t ldo, int64_t M, int64_t N, int64_t K, void* stream) {
  return mq_w4a8_gemm_ws(qa, lda, w4, s_a, s_a_per_row, s_w, bias, x0, w0, out, ldo, M, N, K, nullptr, 0, stream);
}

static int gemm_impl(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                     const float* s_w, const void* bias, const void* x0, const void* w0, void* out, int64_t ldo,
                     int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* stream,
                     bool gated) {
  using namespace mq;
  using namespace mq::gemm;
  if (M == 0) return MQ_OK;  // empty batch: nothing to do (its tensors have null data pointers)
  MQ_REQUIRE(qa && w4 && s_a && s_w && out, "mq_w4a8_gemm: null pointer");
  MQ_REQUIRE(M >= 0 && N > 0 && K > 0, "mq_w4a8_gemm: bad shape (M=%lld N=%lld K=%lld)", (long long)M,
             (long long)N, (long long)K);
  MQ_REQUIRE(K % BK == 0, "mq_w4a8_gemm: K=%lld must be a multiple of %d (zero-pad the codes)", (long long)K, BK);
  MQ_REQUIRE(N % 8 == 0 && ldo % 8 == 0 && ldo >= (gated ? N / 2 : N),
             "mq_w4a8_gemm: N and ldo must be multiples of 8 (N=%lld ldo=%lld)", (long long)N, (long long)ldo);
  MQ_REQUIRE(lda % 16 == 0 && lda >= K, "mq_w4a8_gemm: lda=%lld must be a multiple of 16 and >= K", (long long)lda);
  MQ_REQUIRE(aligned16(qa) && aligned16(w4) && aligned16(out), "mq_w4a8_gemm: pointers must be 16-byte aligned");
  MQ_REQUIRE((x0 == nullptr) == (w0 == nullptr), "mq_w4a8_gemm: x0 and w0 go together");
  MQ_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "mq_w4a8_gemm: dimension exceeds 2^31");
  if (M == 0) return MQ_OK;

  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_w4a8_gemm: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)lda};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)qa, gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("mq_w4a8_gemm: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  Params p;
  p.w4 = w4; p.s_a = s_a; p.s_a_per_row = s_a_per_row; p.s_w = s_w;
  p.bias = (const __half*)bias; p.x0 = (const __half*)x0; p.w0 = (const __half*)w0;
  p.out = (__half*)out; p.ldo = ldo;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.m_tiles = (int)((M + BM - 1) / BM);
  p.n_tiles = (int)((N + BN - 1) / BN);
  p.k_tiles = (int)(K / BK);
  p.dbg = probe_mask();
  // kernel choice: plan_kernel(), unless a test / tool forces one of the two persistent kernels
  int kernel = plan_kernel(M, N, K, workspace != nullptr ? workspace_bytes : 0);
  const int forced = g_force_kernel.load(std::memory_order_relaxed);
  if (forced >= 3 && forced <= 5) kernel = forced;
  if (forced == 2 && workspace != nullptr && splitk_factor(M, N, K) >= 2 && workspace_bytes >= splitk_workspace_bytes(M, N, K))
    kernel = 2;  // (tests: the split-K kernel wherever it is applicable)
  if (p.dbg != 0 && kernel < 3) kernel = 3;  // the probes live in the persistent kernels
  if (gated) kernel = 4;                     // (the caller has checked that this is the planned kernel)
  cudaStream_t st = (cudaStream_t)stream;
  if (kernel == 1) {
    const int rc = launch_w4a8_gemv(qa, lda, p, workspace, st);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemv_kernel");
    return MQ_OK;
  }
  if (kernel == 2) {
    const int rc = launch_w4a8_gemm_sk(tmap, p, workspace, st);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_sk_kernel");
    return MQ_OK;
  }
  if (kernel == 4) {
    // the 256 x 256 kernel writes its output with TMA: 32 x 32 fp16 boxes out of swizzled staging
    CUtensorMap tmap_out;
    const cuuint64_t odim[2] = {(cuuint64_t)(gated ? N / 2 : N), (cuuint64_t)M};
    const cuuint64_t ostride[1] = {(cuuint64_t)ldo * 2};
    const cuuint32_t obox[2] = {32, 32};
    const CUresult cro = encode(&tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, odim, ostride, obox, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cro != CUDA_SUCCESS) {
      set_error("mq_w4a8_gemm: cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)cro);
      return MQ_ERR_CUDA;
    }
    const int rc = launch_w4a8_gemm_v4(tmap, tmap_out, p, st, gated);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_v4_kernel");
    return MQ_OK;
  }
  if (kernel == 5) {
    const int rc = launch_w4a8_gemm_v5(tmap, p, st);
    if (rc != MQ_OK) return rc;
    MQ_LAUNCH_CHECK("w4a8_gemm_v5_kernel");
    return MQ_OK;
  }
  const int rc = launch_w4a8_gemm_v2(tmap, p, st);
  if (rc != MQ_OK) return rc;
  MQ_LAUNCH_CHECK("w4a8_gemm_v2_kernel");
  return MQ_OK;
}

extern "C" int mq_w4a8_gemm_ws(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                               const float* s_w, const void* bias, const void* x0, const void* w0, void* out,
                               int64_t ldo, int64_t M, int64_t N, int64_t K, void* workspace,
                               size_t workspace_bytes, void* stream) {
  return gemm_impl(qa, lda, w4, s_a, s_a_per_row, s_w, bias, x0, w0, out, ldo, M, N, K, workspace, workspace_bytes, stream,
                   false);
}

// ---- gated MLP: a1 * silu(a2) in the epilogue -------------------------------------
namespace mq {
namespace gemm {

// out[m, f] = in[m, 256 (f / 128) + f % 128] * silu(in[m, 256 (f / 128) + 128 + f % 128]): the gate on an
// un-gated GEMM output over interleaved weights -- same arithmetic as the fused epilogue (silu_mul8)
__global__ void __launch_bounds__(256) silu_mul_interleaved_kernel(const __half* __restrict__ in, int64_t ldi,
                                                                  __half* __restrict__ out, int64_t ldo, int64_t M, int F) {
  const int chunks = F / 8;  // 16-byte chunks per output row
  const int64_t total = M * chunks;
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = id / chunks;
    const int f = (int)(id - m * chunks) * 8;
    const __half* src = in + m * ldi + (f >> 7) * 256 + (f & 127);
    const uint4 a1 = *reinterpret_cast<const uint4*>(src), a2 = *reinterpret_cast<const uint4*>(src + 128);
    *reinterpret_cast<uint4*>(out + m * ldo + f) = silu_mul8(a1, a2);
  }
}

}  // namespace gemm
}  // namespace mq

extern "C" int mq_silu_mul_interleaved_f16(const void* in, int64_t ldi, void* out, int64_t ldo, int64_t M, int64_t F,
                                           void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;
  MQ_REQUIRE(in && out, "mq_silu_mul_interleaved_f16: null pointer");
  MQ_REQUIRE(M > 0 && F > 0 && F % 128 == 0 && ldi % 8 == 0 && ldo % 8 == 0 && ldi >= 2 * F && ldo >= F && F < (1ll << 30),
             "mq_silu_mul_interleaved_f16: bad shape (F=%lld must be a multiple of 128)", (long long)F);
  MQ_REQUIRE(aligned16(in) && aligned16(out), "mq_silu_mul_interleaved_f16: pointers must be 16-byte aligned");
  const int64_t total = M * (F / 8);
  const int64_t want = (total + 255) / 256, cap = (int64_t)num_sms() * 16;
  gemm::silu_mul_interleaved_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)in, ldi, (__half*)out, ldo, M, (int)F);
  MQ_LAUNCH_CHECK("mq_silu_mul_interleaved_f16");
  return MQ_OK;
}

extern "C" int mq_w4a8_gemm_gated_ws(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a, int s_a_per_row,
                                     const float* s_w, const void* bias, void* out, int64_t ldo, int64_t M, int64_t N2,
                                     int64_t K, void* tmp, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mq;
  using namespace mq::gemm;
  if (M == 0) return MQ_OK;
  MQ_REQUIRE(N2 > 0 && N2 % 256 == 0, "mq_w4a8_gemm_gated: N2=%lld must be a multiple of 256 (128 channels of w1, 128 of w2)",
             (long long)N2);
  MQ_REQUIRE(K > 0 && K % BK == 0, "mq_w4a8_gemm_gated: K=%lld must be a multiple of %d", (long long)K, BK);
  int kernel = plan_kernel(M, N2, K, workspace != nullptr ? workspace_bytes : 0);
  const int forced = g_force_kernel.load(std::memory_order_relaxed);
  if (forced >= 3 && forced <= 5) kernel = forced;
  if (kernel == 4)
    return gemm_impl(qa, lda, w4, s_a, s_a_per_row, s_w, bias, nullptr, nullptr, out, ldo, M, N2, K, workspace,
                     workspace_bytes, stream, true);
  // shapes the 256 x 256 kernel does not serve: the plain GEMM into tmp, then the same gate arithmetic
  MQ_REQUIRE(tmp != nullptr, "mq_w4a8_gemm_gated: M=%lld runs the two-step path and needs tmp [M, N2] fp16", (long long)M);
  const int rc = gemm_impl(qa, lda, w4, s_a, s_a_per_row, s_w, bias, nullptr, nullptr, tmp, N2, M, N2, K, workspace,
                           workspace_bytes, stream, false);
  if (rc != MQ_OK) return rc;
  return mq_silu_mul_interleaved_f16(tmp, N2, out, ldo, M, N2 / 2, stream);
}

extern "C" int mq_debug_force_gemm_kernel(int kernel) {
  return mq::gemm::g_force_kernel.exchange(kernel >= 2 && kernel <= 5 ? kernel : 0);
}
