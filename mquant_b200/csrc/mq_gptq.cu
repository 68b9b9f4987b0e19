// GPTQ acceleration.
//
//   mq_gptq_hessian_update   GPTQ.add_batch (fake_quant/gptq/gptq_utils.py:182-195):
//                              H = H * n/(n+t) + (2/(n+t)) * X^T X          X: [T, C] fp16, H: [C, C] fp32
//                            on tcgen05.  The activations are fp16 VALUES, so every product x_i x_j is
//                            exact in fp32 and the tensor core (fp16 x fp16 -> fp32) computes what the
//                            reference's fp32 matmul computes, up to the order of the fp32 additions.
//   mq_gptq_block_f32        the inner loop of GPTQ.fasterquant (:254-283): for the columns of one
//                            block, quantise column i, feed its error into the columns right of it.
//                            Rows are independent: one thread per row, the block's slice of Hinv in
//                            shared memory; every operation is the reference's fp32 operation in the
//                            reference's order (bit-exact against the eager loop on the same Hinv).
//
// Hessian kernel: one CTA per 128 x 256 tile of H; X is the MN-major operand on both sides (A =
// X^T: 64-column blocks of 64 tokens, B the same), 64 tokens per pipeline stage, 4 stages of 48 KB;
// warp 0 TMA, warp 1 MMA issuer, warps 2-9 epilogue.
//
// Precision: the tensor core adds the exact products of a K = 16 step into its fp32 accumulator with
// truncation -- measured ~9e-8 of the running sum lost per step, a BIAS that grows linearly with the
// number of tokens (2.3e-5 of the diagonal after 4096 tokens, 40x the error of an fp32 FFMA chain).
// So a TMEM accumulator only ever holds CHUNK_KB * 64 tokens: two accumulators (2 x 256 columns)
// alternate, the eight epilogue warps (thread = one row of the tile, 128 of its columns) drain the
// finished one into fp32 registers with round-to-nearest adds while the other one is being filled.
// The error is then that of the reference's own fp32 matmul (tests: <= 4e-6 of max |H|).
#include <cuda.h>

#include "mq_common.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace gptq {
using namespace mq::tc;

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int CHUNK_KB = 2;  // k-blocks (of 64 tokens) per TMEM accumulation chain
constexpr int A_BYTES = (BM / 64) * BK * 128, B_BYTES = (BN / 64) * BK * 128, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int HESS_THREADS = 320, NBAR = 2 * STAGES + 4;

__global__ void __launch_bounds__(HESS_THREADS, 1)
hessian_kernel(const __grid_constant__ CUtensorMap tmap_x, float* __restrict__ H, int C, int64_t T, float alpha, float beta) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + STAGES * STAGE_BYTES + 8 * NBAR);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int nkb = (int)((T + BK - 1) / BK);
  const int nchunks = (nkb + CHUNK_KB - 1) / CHUNK_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 8);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(full(s), STAGE_BYTES);
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tmap_x, i0 + 64 * c, kb * BK, full(s));
#pragma unroll
        for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tmap_x, j0 + 64 * c, kb * BK, full(s));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // F16 x F16 -> F32, A and B MN-major, M = 128, N = 256
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int b = c & 1;
      mbar_wait(acc_empty(b), (uint32_t)(((c >> 1) & 1) ^ 1));  // drained by the epilogue (free at first use)
      tc_fence_after();
      const int kb_end = min(nkb, (c + 1) * CHUNK_KB);
      for (int kb = c * CHUNK_KB; kb < kb_end; ++kb) {
        mbar_wait(full(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t ad = make_sw128_desc(sa, BK * 128, 1024), bd = make_sw128_desc(sb, BK * 128, 1024);
          const bool first = kb == c * CHUNK_KB;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)  // 16 tokens = 2048 bytes further down the 64-column blocks
            umma_f16(tmem_base + (uint32_t)(b * BN), ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc,
                     (first && ks == 0) ? 0u : 1u);
          umma_commit(empty(s));
          if (kb == kb_end - 1) umma_commit(acc_full(b));
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // epilogue: lane = row i of the tile (TMEM lane quadrant = warp % 4), 128 columns per thread
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int i = i0 + quad * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
    float sum[128];
#pragma unroll
    for (int e = 0; e < 128; ++e) sum[e] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int b = c & 1;
      mbar_wait_sleep(acc_full(b), (uint32_t)((c >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(b * BN + cc * 32), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) sum[cc * 32 + e] = __fadd_rn(sum[cc * 32 + e], __uint_as_float(v[e]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(b));
    }
    if (i < C) {
      float* hrow = H + (int64_t)i * C;
      const int jb = j0 + half * 128;
      if (jb + 128 <= C && (C & 3) == 0) {
#pragma unroll
        for (int e = 0; e < 128; e += 4) {
          float4 h = *reinterpret_cast<const float4*>(hrow + jb + e);
          h.x = __fmaf_rn(beta, sum[e], __fmul_rn(alpha, h.x));
          h.y = __fmaf_rn(beta, sum[e + 1], __fmul_rn(alpha, h.y));
          h.z = __fmaf_rn(beta, sum[e + 2], __fmul_rn(alpha, h.z));
          h.w = __fmaf_rn(beta, sum[e + 3], __fmul_rn(alpha, h.w));
          *reinterpret_cast<float4*>(hrow + jb + e) = h;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 128; ++e)
          if (jb + e < C) hrow[jb + e] = __fmaf_rn(beta, sum[e], __fmul_rn(alpha, hrow[jb + e]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- fasterquant inner loop ----------------------------------------------------------------------
// W: [rows, ldw] fp32, the block's columns start at W (already offset); Hinv1: [count, ldh] fp32 (upper
// triangle of the block); scale: [rows]; outputs Q / Err / Losses: [rows, count].
constexpr int GB_ROWS = 128;  // rows (threads) per CTA
constexpr int GB_MAXC = 128;  // columns per block

__global__ void __launch_bounds__(GB_ROWS)
gptq_block_kernel(float* __restrict__ W, int64_t ldw, const float* __restrict__ Hinv1, int64_t ldh,
                  const float* __restrict__ scale, float maxq, float* __restrict__ Q, float* __restrict__ Err,
                  float* __restrict__ Losses, int rows, int count) {
  extern __shared__ float sm[];
  float* sh = sm;                          // [count][count] Hinv1
  float* sw = sm + GB_MAXC * GB_MAXC;      // [GB_ROWS][count + 1] the rows of this CTA
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * GB_ROWS;
  const int wst = count + 1;
  for (int id = tid; id < count * count; id += GB_ROWS) sh[id] = Hinv1[(int64_t)(id / count) * ldh + (id % count)];
  for (int id = tid; id < GB_ROWS * count; id += GB_ROWS) {  // coalesced: consecutive threads, consecutive columns
    const int r = id / count, cidx = id % count;
    sw[r * wst + cidx] = (r0 + r < rows) ? W[(int64_t)(r0 + r) * ldw + cidx] : 0.f;
  }
  __syncthreads();
  const int r = r0 + tid;
  if (r < rows) {
    float* w = sw + tid * wst;
    const float s = scale[r];
    const float lo = -(maxq + 1.f);
    for (int i = 0; i < count; ++i) {
      const float wi = w[i];
      const float d = sh[i * count + i];
      // sym_quant_dequant (quant_utils.py:42-57): scale * clamp(round(x / scale), -(maxq+1), maxq)
      const float q = __fmul_rn(s, fminf(fmaxf(rintf(__fdiv_rn(wi, s)), lo), maxq));
      const float diff = __fsub_rn(wi, q);
      Q[(int64_t)r * count + i] = q;
      Losses[(int64_t)r * count + i] = __fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(d, d));
      const float e = __fdiv_rn(diff, d);
      Err[(int64_t)r * count + i] = e;
      // W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0)): a rounded product, then a rounded difference
      for (int j = i; j < count; ++j) w[j] = __fsub_rn(w[j], __fmul_rn(e, sh[i * count + j]));
    }
  }
  __syncthreads();
  for (int id = tid; id < GB_ROWS * count; id += GB_ROWS) {
    const int rr = id / count, cidx = id % count;
    if (r0 + rr < rows) W[(int64_t)(r0 + rr) * ldw + cidx] = sw[rr * wst + cidx];
  }
}

}  // namespace gptq
}  // namespace mq

extern "C" int mq_gptq_hessian_update(const void* x, int64_t ldx, int64_t T, int C, float* H, float alpha, float beta,
                                      void* stream) {
  using namespace mq;
  using namespace mq::gptq;
  if (T == 0) return MQ_OK;
  MQ_REQUIRE(x != nullptr && H != nullptr, "mq_gptq_hessian_update: null pointer");
  MQ_REQUIRE(T > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C, "mq_gptq_hessian_update: bad shape (C=%d ldx=%lld)", C,
             (long long)ldx);
  MQ_REQUIRE(aligned16(x) && aligned16(H), "mq_gptq_hessian_update: pointers must be 16-byte aligned");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_gptq_hessian_update: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  CUtensorMap tmap;  // [T][C]: tokens / columns outside are zero filled and add nothing
  const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)T};
  const cuuint64_t gstride[1] = {(cuuint64_t)ldx * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)BK};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("mq_gptq_hessian_update: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  const int smem = STAGES * STAGE_BYTES + 8 * NBAR + 16 + 1024;
  MQ_DYN_SMEM(hessian_kernel, smem);
  dim3 grid((C + BN - 1) / BN, (C + BM - 1) / BM);
  hessian_kernel<<<grid, HESS_THREADS, smem, (cudaStream_t)stream>>>(tmap, H, C, T, alpha, beta);
  MQ_LAUNCH_CHECK("mq_gptq_hessian_update");
  return MQ_OK;
}

extern "C" int mq_gptq_block_f32(float* W, int64_t ldw, const float* Hinv1, int64_t ldh, const float* scale, float maxq,
                                 float* Q, float* Err, float* Losses, int rows, int count, void* stream) {
  using namespace mq;
  using namespace mq::gptq;
  if (rows == 0 || count == 0) return MQ_OK;
  MQ_REQUIRE(W != nullptr && Hinv1 != nullptr && scale != nullptr && Q != nullptr && Err != nullptr && Losses != nullptr,
             "mq_gptq_block_f32: null pointer");
  MQ_REQUIRE(rows > 0 && count > 0 && count <= GB_MAXC && ldw >= count && ldh >= count,
             "mq_gptq_block_f32: bad shape (rows=%d count=%d, at most %d columns per block)", rows, count, GB_MAXC);
  const int smem = (GB_MAXC * GB_MAXC + GB_ROWS * (GB_MAXC + 1)) * (int)sizeof(float);
  MQ_DYN_SMEM(gptq_block_kernel, smem);
  gptq_block_kernel<<<(rows + GB_ROWS - 1) / GB_ROWS, GB_ROWS, smem, (cudaStream_t)stream>>>(W, ldw, Hinv1, ldh, scale, maxq, Q,
                                                                                             Err, Losses, rows, count);
  MQ_LAUNCH_CHECK("mq_gptq_block_f32");
  return MQ_OK;
}
