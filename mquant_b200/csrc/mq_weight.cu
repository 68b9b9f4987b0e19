// Weight-side kernels: per-channel RTN quantisation, INT4 packing, and the CUDA-core
// cross-check GEMM.  All offline / test-only: correctness over speed.
#include "mq_common.cuh"
#include "mq_w4_layout.cuh"

namespace mq {

// WeightQuantizer.find_params (sym, perchannel; quant_utils.py:446-466) + sym_quant (:42-45).
// One warp per output channel.
__global__ void __launch_bounds__(256)
weight_quant_rtn_kernel(const __half* __restrict__ w, int64_t ldw, int8_t* __restrict__ codes,
                        float* __restrict__ s_w, const float* __restrict__ scales_in, int64_t N, int K,
                        int bits) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const __half* row = w + n * ldw;
  const float maxq = (float)((1 << (bits - 1)) - 1);
  float s;
  if (scales_in != nullptr) {
    s = scales_in[n];
  } else {
    float mx = 0.f, mn = 0.f;  // minimum(x.min, 0), maximum(x.max, 0) in fp32 (:458-460)
    for (int k = lane; k < K; k += 32) {
      const float v = __half2float(row[k]);
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    const float xmax = fmaxf(fmaxf(fabsf(mn), mx), 1e-5f);  // .clamp(min=1e-5)
    s = __fdiv_rn(xmax, maxq);
  }
  if (lane == 0) s_w[n] = s;
  for (int k = lane; k < K; k += 32) {
    const float q = quant_code_exact(__half2float(row[k]), s, -(maxq + 1.f), maxq);
    codes[n * (int64_t)K + k] = (int8_t)__float2int_rn(q);
  }
}

// one thread per 16-byte packed chunk
__global__ void __launch_bounds__(256)
pack_w4_kernel(const int8_t* __restrict__ codes, int64_t ldc, uint8_t* __restrict__ w4, int64_t N, int64_t K,
               unsigned flags) {
  const int64_t k_tiles = K / kW4TileK;
  const int64_t n_tiles = (N + kW4TileN - 1) / kW4TileN;
  const int64_t total = n_tiles * k_tiles * (kW4BlobBytes / 16);
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t blob = id / (kW4BlobBytes / 16);
    const int q = (int)(id - blob * (kW4BlobBytes / 16));
    const int j = q / kW4TileN, r = q % kW4TileN;
    const int64_t nb = blob / k_tiles, kb = blob - nb * k_tiles;
    const int64_t n = nb * kW4TileN + r;
    const int64_t k0 = kb * kW4TileK + 32 * j;
    uint8_t out[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      int lo = 0, hi = 0;
      if (n < N) {
        lo = codes[n * ldc + k0 + b];
        hi = codes[n * ldc + k0 + 16 + b];
        if ((flags & MQ_FLAG_SPLIT) && k0 + b == 0) lo = 0;
      }
      out[b] = (uint8_t)((lo & 0xf) | ((hi & 0xf) << 4));
    }
    *reinterpret_cast<uint4*>(w4 + id * 16) = *reinterpret_cast<const uint4*>(out);
  }
}

// reference interchange format (pack_i4, quant_utils.py:61-68): byte i of a row holds
// code[2i] (low nibble) and code[2i+1] (high nibble)
__global__ void __launch_bounds__(256)
repack_w4_kernel(const uint8_t* __restrict__ ref, uint8_t* __restrict__ w4, int64_t N, int64_t K) {
  const int64_t k_tiles = K / kW4TileK;
  const int64_t n_tiles = (N + kW4TileN - 1) / kW4TileN;
  const int64_t total = n_tiles * k_tiles * (kW4BlobBytes / 16);
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t blob = id / (kW4BlobBytes / 16);
    const int q = (int)(id - blob * (kW4BlobBytes / 16));
    const int j = q / kW4TileN, r = q % kW4TileN;
    const int64_t nb = blob / k_tiles, kb = blob - nb * k_tiles;
    const int64_t n = nb * kW4TileN + r;
    const int64_t k0 = kb * kW4TileK + 32 * j;
    uint8_t out[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      unsigned lo = 0, hi = 0;
      if (n < N) {
        const int64_t ka = k0 + b, kb2 = k0 + 16 + b;
        const uint8_t pa = ref[n * (K / 2) + (ka >> 1)], pb = ref[n * (K / 2) + (kb2 >> 1)];
        lo = (ka & 1) ? (pa >> 4) : (pa & 0xf);
        hi = (kb2 & 1) ? (pb >> 4) : (pb & 0xf);
      }
      out[b] = (uint8_t)(lo | (hi << 4));
    }
    *reinterpret_cast<uint4*>(w4 + id * 16) = *reinterpret_cast<const uint4*>(out);
  }
}

__device__ __forceinline__ int w4_code_at(const uint8_t* __restrict__ w4, int64_t n, int64_t k, int64_t k_tiles) {
  const int64_t nb = n / kW4TileN, kb = k / kW4TileK;
  const int r = (int)(n - nb * kW4TileN);
  const int kk = (int)(k - kb * kW4TileK);
  const int j = kk >> 5, b = kk & 15, hi = (kk >> 4) & 1;
  const uint8_t byte = w4[w4_blob_offset(nb, kb, k_tiles) + w4_chunk_offset(j, r) + b];
  const int nib = hi ? (byte >> 4) : (byte & 0xf);
  return nib >= 8 ? nib - 16 : nib;
}

__global__ void __launch_bounds__(256)
unpack_w4_kernel(const uint8_t* __restrict__ w4, int8_t* __restrict__ codes, int64_t ldc, int64_t N, int64_t K) {
  const int64_t k_tiles = K / kW4TileK;
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < N * K;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = id / K, k = id - n * K;
    codes[n * ldc + k] = (int8_t)w4_code_at(w4, n, k, k_tiles);
  }
}

// CUDA-core cross-check of mq_w4a8_gemm's contract: one thread per output element.
__global__ void __launch_bounds__(256)
w4a8_gemm_simt_kernel(const int8_t* __restrict__ qa, int64_t lda, const uint8_t* __restrict__ w4,
                      const float* __restrict__ s_a, int s_a_per_row, const float* __restrict__ s_w,
                      const __half* __restrict__ bias, const __half* __restrict__ x0,
                      const __half* __restrict__ w0, __half* __restrict__ out, int64_t ldo, int64_t M, int64_t N,
                      int64_t K) {
  const int64_t k_tiles = K / kW4TileK;
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < M * N;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = id / N, n = id - m * N;
    const int64_t nb = n / kW4TileN;
    const int r = (int)(n - nb * kW4TileN);
    const int8_t* arow = qa + m * lda;
    int acc = 0;
    for (int64_t kb = 0; kb < k_tiles; ++kb) {
      const uint8_t* blob = w4 + w4_blob_offset(nb, kb, k_tiles);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 c = *reinterpret_cast<const uint4*>(blob + w4_chunk_offset(j, r));
        const uint8_t* cb = reinterpret_cast<const uint8_t*>(&c);
        const int8_t* ap = arow + kb * kW4TileK + 32 * j;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          int lo = cb[b] & 0xf, hi = cb[b] >> 4;
          lo = lo >= 8 ? lo - 16 : lo;
          hi = hi >= 8 ? hi - 16 : hi;
          acc += lo * (int)ap[b] + hi * (int)ap[16 + b];
        }
      }
    }
    const float sa = s_a_per_row ? s_a[m] : s_a[0];
    float y = __fmul_rn((float)acc, __fmul_rn(sa, s_w[n]));
    if (x0 != nullptr) y = __fadd_rn(y, __fmul_rn(__half2float(x0[m]), __half2float(w0[n])));
    if (bias != nullptr) y = __fadd_rn(y, __half2float(bias[n]));
    out[m * ldo + n] = __float2half_rn(y);
  }
}

static int simple_grid(int64_t items) {
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mq

extern "C" int mq_weight_quant_rtn(const void* w_f16, int64_t ldw, int8_t* codes, float* s_w,
                                   const float* scales_in, int64_t N, int K, int bits, void* stream) {
  using namespace mq;
  MQ_REQUIRE(w_f16 && codes && s_w, "mq_weight_quant_rtn: null pointer");
  MQ_REQUIRE(N > 0 && K > 0 && ldw >= K && bits >= 2 && bits <= 8, "mq_weight_quant_rtn: bad shape/bits (N=%lld K=%d bits=%d)",
             (long long)N, K, bits);
  const int64_t blocks = (N * 32 + 255) / 256;
  weight_quant_rtn_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)w_f16, ldw, codes, s_w,
                                                                               scales_in, N, K, bits);
  MQ_LAUNCH_CHECK("weight_quant_rtn_kernel");
  return MQ_OK;
}

extern "C" size_t mq_w4_packed_bytes(int64_t N, int64_t K) {
  if (N <= 0 || K <= 0 || K % mq::kW4TileK != 0) return 0;
  const int64_t n_tiles = (N + mq::kW4TileN - 1) / mq::kW4TileN;
  return (size_t)(n_tiles * (K / mq::kW4TileK) * mq::kW4BlobBytes);
}

extern "C" int mq_pack_w4(const int8_t* codes, int64_t ldc, uint8_t* w4, int64_t N, int64_t K, unsigned flags,
                          void* stream) {
  using namespace mq;
  MQ_REQUIRE(codes && w4, "mq_pack_w4: null pointer");
  MQ_REQUIRE(N > 0 && K > 0 && K % kW4TileK == 0 && ldc >= K, "mq_pack_w4: K=%lld must be a multiple of %d",
             (long long)K, kW4TileK);
  MQ_REQUIRE(aligned16(w4), "mq_pack_w4: w4 must be 16-byte aligned");
  pack_w4_kernel<<<simple_grid((int64_t)mq_w4_packed_bytes(N, K) / 16), 256, 0, (cudaStream_t)stream>>>(
      codes, ldc, w4, N, K, flags);
  MQ_LAUNCH_CHECK("pack_w4_kernel");
  return MQ_OK;
}

extern "C" int mq_repack_w4_from_ref(const uint8_t* ref_packed, uint8_t* w4, int64_t N, int64_t K, void* stream) {
  using namespace mq;
  MQ_REQUIRE(ref_packed && w4, "mq_repack_w4_from_ref: null pointer");
  MQ_REQUIRE(N > 0 && K > 0 && K % kW4TileK == 0, "mq_repack_w4_from_ref: K=%lld must be a multiple of %d",
             (long long)K, kW4TileK);
  MQ_REQUIRE(aligned16(w4), "mq_repack_w4_from_ref: w4 must be 16-byte aligned");
  repack_w4_kernel<<<simple_grid((int64_t)mq_w4_packed_bytes(N, K) / 16), 256, 0, (cudaStream_t)stream>>>(
      ref_packed, w4, N, K);
  MQ_LAUNCH_CHECK("repack_w4_kernel");
  return MQ_OK;
}

extern "C" int mq_unpack_w4(const uint8_t* w4, int8_t* codes, int64_t ldc, int64_t N, int64_t K, void* stream) {
  using namespace mq;
  MQ_REQUIRE(codes && w4, "mq_unpack_w4: null pointer");
  MQ_REQUIRE(N > 0 && K > 0 && K % kW4TileK == 0 && ldc >= K, "mq_unpack_w4: K=%lld must be a multiple of %d",
             (long long)K, kW4TileK);
  unpack_w4_kernel<<<simple_grid(N * K), 256, 0, (cudaStream_t)stream>>>(w4, codes, ldc, N, K);
  MQ_LAUNCH_CHECK("unpack_w4_kernel");
  return MQ_OK;
}

extern "C" int mq_w4a8_gemm_simt(const int8_t* qa, int64_t lda, const uint8_t* w4, const float* s_a,
                                 int s_a_per_row, const float* s_w, const void* bias, const void* x0,
                                 const void* w0, void* out, int64_t ldo, int64_t M, int64_t N, int64_t K,
                                 void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;  // empty batch
  MQ_REQUIRE(qa && w4 && s_a && s_w && out, "mq_w4a8_gemm_simt: null pointer");
  MQ_REQUIRE(M >= 0 && N > 0 && K > 0 && K % kW4TileK == 0 && lda >= K && ldo >= N,
             "mq_w4a8_gemm_simt: bad shape (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
  MQ_REQUIRE((x0 == nullptr) == (w0 == nullptr), "mq_w4a8_gemm_simt: x0 and w0 go together");
  if (M == 0) return MQ_OK;
  w4a8_gemm_simt_kernel<<<simple_grid(M * N), 256, 0, (cudaStream_t)stream>>>(
      qa, lda, w4, s_a, s_a_per_row, s_w, (const __half*)bias, (const __half*)x0, (const __half*)w0,
      (__half*)out, ldo, M, N, K);
  MQ_LAUNCH_CHECK("w4a8_gemm_simt_kernel");
  return MQ_OK;
}
