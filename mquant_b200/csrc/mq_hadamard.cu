// Online Walsh-Hadamard rotation fused with the static INT8 quantiser.
//
//   y = x . (hadK (x) H_P)^T / sqrt(n),  n = had_k * P,  P a power of two
//   (matmul_hadU_cuda, fake_quant/hadamard_utils.py:115-128), then UniformQuantizer.quant.
//
// Rounding points mirror the reference: the FHT factor runs in fp32 and is rounded to fp16
// (fast_hadamard_transform returns x.dtype), the hadK factor accumulates in fp32 and is rounded
// to fp16 (hadK.to(fp16) @ input), the quantiser then divides float(fp16 value) by the scale.
// With MQ_FLAG_FP32_HAD the intermediate rounding is skipped (quant_utils.py:336-339).
//
// Kernels in this file
//   fwht_warp_kernel   n = 64..2048 (had_k = 1): one (sub-)warp per row, registers + shuffles
//   fwht_cta_kernel    n = 4096 / 8192 (had_k = 1): 4 warps per row, one shared-memory exchange
//   hadk_simt_kernel   had_k > 1: FHT_P per segment in registers, hadK factor on CUDA cores from
//                      shared memory.  Correctness baseline and fallback (fp16 output for
//                      calibration, --fp32_had, n/had_k = 256/512); the hot path (codes only,
//                      n/had_k = 64 or 128) runs mq_hadamard_tc.cu on tcgen05.
//
// HBM traffic per token: 2*n_in read + n written (codes) [+ 2*n if y_f16 is requested].
#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"

namespace mq {

// in-register butterflies over the 8 elements of each chunk (index bits 0..2) and across the
// R chunks a thread owns (the top log2(R) index bits)
template <int R>
__device__ __forceinline__ void butterflies_in_thread(float (&v)[R][8]) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
#pragma unroll
    for (int s = 1; s < 8; s <<= 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if ((e & s) == 0) {
          const float a = v[j][e], b = v[j][e + s];
          v[j][e] = a + b;
          v[j][e + s] = a - b;
        }
      }
    }
  }
#pragma unroll
  for (int s = 1; s < R; s <<= 1) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if ((j & s) == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a = v[j][e], b = v[j + s][e];
          v[j][e] = a + b;
          v[j + s][e] = a - b;
        }
      }
    }
  }
}

// butterflies across `LANES` consecutive lanes (index bits carried by the lane id)
template <int R, int LANES>
__device__ __forceinline__ void butterflies_across_lanes(float (&v)[R][8], int lane) {
#pragma unroll
  for (int s = 1; s < LANES; s <<= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float o = __shfl_xor_sync(0xffffffffu, v[j][e], s);
        v[j][e] = upper ? (o - v[j][e]) : (v[j][e] + o);
      }
    }
  }
}

// load 8 halfs of a zero-padded row
__device__ __forceinline__ void load_chunk(const __half* row, int col, int n_in, float (&f)[8]) {
  if (col + 8 <= n_in) {
    unpack8(ldg_stream(row + col), f);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
  }
}

// finish 8 consecutive outputs: scale, fp16 rounding point, then y_f16 and/or codes
__device__ __forceinline__ void emit8(const HadArgs& a, int64_t m, int col, float (&f)[8], float s_act,
                                      bool scale_here) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float t = scale_here ? __fmul_rn(f[e], a.norm) : f[e];
    f[e] = round_f16(t);
  }
  if (a.y != nullptr) stg_stream(a.y + m * (int64_t)a.n + col, pack8_f16(f));
  if (a.q != nullptr) {
    uint2 r = quant8(f, make_qscale(s_act));
    if (col == 0 && (a.flags & MQ_FLAG_SPLIT)) {
      a.x0_out[m] = __float2half_rn(f[0]);
      r.x &= 0xffffff00u;
    }
    stg_stream8(a.q + m * (int64_t)a.n + col, r);
  }
}

__device__ __forceinline__ float row_scale(const HadArgs& a, int64_t m) {
  if (a.q == nullptr) return 1.f;
  const float s = (a.modality != nullptr && a.modality[m] != 0) ? __ldg(a.scales + 1) : __ldg(a.scales);
  return s;
}

// ---------------------------------------------------------------------------------------------
// n <= 2048: LANES = min(32, n/8) lanes per row, R = n / (8*LANES) chunks per lane.
// chunk c of lane l, slot j:  c = l + LANES*j  -> loads and stores are fully coalesced.
// ---------------------------------------------------------------------------------------------
template <int LOG_N>
__global__ void __launch_bounds__(128) fwht_warp_kernel(const HadArgs a) {
  constexpr int N = 1 << LOG_N;
  constexpr int LANES = (N / 8) < 32 ? (N / 8) : 32;
  constexpr int R = N / (8 * LANES);
  constexpr int ROWS_PER_WARP = 32 / LANES;
  const int lane_in_warp = threadIdx.x & 31;
  const int lane = lane_in_warp % LANES;
  const int sub = lane_in_warp / LANES;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // all lanes of a warp iterate together (shuffles are warp-wide); rows past M are masked
  for (int64_t base = warp * ROWS_PER_WARP; base < a.M; base += nwarps * ROWS_PER_WARP) {
    const int64_t m = base + sub;
    const bool live = m < a.M;
    float v[R][8];
    const __half* row = a.x + (live ? m : 0) * a.ldx;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (live) load_chunk(row, 8 * (lane + LANES * j), a.n_in, v[j]);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
      }
    }
    butterflies_in_thread<R>(v);
    butterflies_across_lanes<R, LANES>(v, lane);
    if (live) {
      const float s_act = row_scale(a, m);
      if (lane == 0 && a.srow_out != nullptr) a.srow_out[m] = s_act;
#pragma unroll
      for (int j = 0; j < R; ++j) emit8(a, m, 8 * (lane + LANES * j), v[j], s_act, true);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// n = 4096 / 8192: one row per 128-thread CTA iteration.  R = n/1024 chunks per thread,
// chunk c = t + 128*j.  Index bits: e -> 0..2, lane -> 3..7, warp -> 8..9, j -> 10...
// The warp bits are resolved by one exchange through shared memory (fp32, 16-byte pad per
// 128 bytes so that both the 32-byte-strided writes and reads are bank-conflict free).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pad_idx(int i) { return i + ((i >> 5) << 2); }

template <int LOG_N>
__global__ void __launch_bounds__(128) fwht_cta_kernel(const HadArgs a) {
  constexpr int N = 1 << LOG_N;
  constexpr int R = N / 1024;
  static_assert(R >= 4, "cta kernel is for n >= 4096");
  constexpr int JPW = R / 4;  // chunk slots each warp finishes after the exchange
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  const int lane = t & 31, w = t >> 5;
  for (int64_t m = blockIdx.x; m < a.M; m += gridDim.x) {
    float v[R][8];
    const __half* row = a.x + m * a.ldx;
#pragma unroll
    for (int j = 0; j < R; ++j) load_chunk(row, 8 * (t + 128 * j), a.n_in, v[j]);
    butterflies_in_thread<R>(v);
    butterflies_across_lanes<R, 32>(v, lane);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float4* dst = reinterpret_cast<float4*>(sm + pad_idx(8 * (t + 128 * j)));
      dst[0] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
      dst[1] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
    }
    __syncthreads();
    // thread (w, lane) now owns chunk slots j = w*JPW .. w*JPW+JPW-1 for all 4 warp-bit values
    float u[JPW][4][8];
#pragma unroll
    for (int jj = 0; jj < JPW; ++jj) {
#pragma unroll
      for (int wb = 0; wb < 4; ++wb) {
        const int c = lane + 32 * wb + 128 * (w * JPW + jj);
        const float4* src = reinterpret_cast<const float4*>(sm + pad_idx(8 * c));
        const float4 p = src[0], r = src[1];
        u[jj][wb][0] = p.x; u[jj][wb][1] = p.y; u[jj][wb][2] = p.z; u[jj][wb][3] = p.w;
        u[jj][wb][4] = r.x; u[jj][wb][5] = r.y; u[jj][wb][6] = r.z; u[jj][wb][7] = r.w;
      }
    }
    __syncthreads();  // smem is rewritten by the next row
    const float s_act = row_scale(a, m);
    if (t == 0 && a.srow_out != nullptr) a.srow_out[m] = s_act;
#pragma unroll
    for (int jj = 0; jj < JPW; ++jj) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a0 = u[jj][0][e], a1 = u[jj][1][e], a2 = u[jj][2][e], a3 = u[jj][3][e];
        const float b0 = a0 + a1, b1 = a0 - a1, b2 = a2 + a3, b3 = a2 - a3;
        u[jj][0][e] = b0 + b2;
        u[jj][1][e] = b1 + b3;
        u[jj][2][e] = b0 - b2;
        u[jj][3][e] = b1 - b3;
      }
#pragma unroll
      for (int wb = 0; wb < 4; ++wb) {
        const int c = lane + 32 * wb + 128 * (w * JPW + jj);
        emit8(a, m, 8 * c, u[jj][wb], s_act, true);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// had_k > 1 baseline: one CTA (256 threads) per row.
//   stage 1: segment k (P contiguous elements) -> FHT_P in registers (LANES = P/8 lanes, or 32
//            lanes x 2 chunks for P = 512), * 1/sqrt(n), fp16 rounding, t[k][j] to smem (fp16)
//   stage 2: y[k'][j] = sum_k hadK[k'][k] * t[k][j], fp32 accumulation in k order
// ---------------------------------------------------------------------------------------------
template <int LOG_P>
__global__ void __launch_bounds__(256) hadk_simt_kernel(const HadArgs a) {
  constexpr int P = 1 << LOG_P;
  constexpr int LANES = (P / 8) < 32 ? (P / 8) : 32;
  constexpr int R = P / (8 * LANES);
  constexpr int SEG_PER_WARP = 32 / LANES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // t as fp16 (reference rounding) or fp32 (MQ_FLAG_FP32_HAD)
  const bool fp32_had = (a.flags & MQ_FLAG_FP32_HAD) != 0;
  __half* th = reinterpret_cast<__half*>(smem_raw);
  float* tf = reinterpret_cast<float*>(smem_raw);
  const int K = a.had_k;
  const int lane_in_warp = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lane = lane_in_warp % LANES, sub = lane_in_warp / LANES;
  for (int64_t m = blockIdx.x; m < a.M; m += gridDim.x) {
    const __half* row = a.x + m * a.ldx;
    for (int k0 = warp * SEG_PER_WARP; k0 < K; k0 += 8 * SEG_PER_WARP) {
      const int k = k0 + sub;
      const bool live = k < K;
      float v[R][8];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (live) load_chunk(row, k * P + 8 * (lane + LANES * j), a.n_in, v[j]);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
        }
      }
      butterflies_in_thread<R>(v);
      butterflies_across_lanes<R, LANES>(v, lane);
      if (live) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int col = 8 * (lane + LANES * j);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[j][e] = __fmul_rn(v[j][e], a.norm);
          if (fp32_had) {
            float4* d = reinterpret_cast<float4*>(tf + k * P + col);
            d[0] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
            d[1] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
          } else {
            *reinterpret_cast<uint4*>(th + k * P + col) = pack8_f16(v[j]);
          }
        }
      }
    }
    __syncthreads();
    const float s_act = row_scale(a, m);
    if (threadIdx.x == 0 && a.srow_out != nullptr) a.srow_out[m] = s_act;
    const int chunks = K * (P / 8);
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
      const int kp = c / (P / 8);
      const int col = (c - kp * (P / 8)) * 8;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      const __half* hrow = a.hadk + kp * K;
      for (int k = 0; k < K; ++k) {
        const float h = __half2float(__ldg(hrow + k));
        float f[8];
        if (fp32_had) {
          const float4* s = reinterpret_cast<const float4*>(tf + k * P + col);
          const float4 p = s[0], r = s[1];
          f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = p.w; f[4] = r.x; f[5] = r.y; f[6] = r.z; f[7] = r.w;
        } else {
          unpack8(*reinterpret_cast<const uint4*>(th + k * P + col), f);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(h, f[e], acc[e]);
      }
      emit8(a, m, kp * P + col, acc, s_act, false);
    }
    __syncthreads();
  }
}

template <int LOG_N>
static int launch_warp(const HadArgs& a, cudaStream_t st) {
  constexpr int N = 1 << LOG_N;
  constexpr int LANES = (N / 8) < 32 ? (N / 8) : 32;
  const int64_t warps = (a.M + (32 / LANES) - 1) / (32 / LANES);
  int64_t blocks = (warps + 3) / 4;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  fwht_warp_kernel<LOG_N><<<(int)blocks, 128, 0, st>>>(a);
  return MQ_OK;
}

template <int LOG_N>
static int launch_cta(const HadArgs& a, cudaStream_t st) {
  constexpr int N = 1 << LOG_N;
  const size_t smem = (size_t)(N + N / 8) * sizeof(float);
  MQ_DYN_SMEM(fwht_cta_kernel<LOG_N>, smem);
  int64_t blocks = a.M;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  fwht_cta_kernel<LOG_N><<<(int)blocks, 128, smem, st>>>(a);
  return MQ_OK;
}

template <int LOG_P>
static int launch_hadk(const HadArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)a.n * ((a.flags & MQ_FLAG_FP32_HAD) ? 4 : 2);
  MQ_REQUIRE(smem <= 200 * 1024, "mq_hadamard_quant_i8: n=%d too large for the had_k>1 kernel", a.n);
  MQ_DYN_SMEM(hadk_simt_kernel<LOG_P>, smem);
  int64_t blocks = a.M;
  const int64_t cap = (int64_t)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  hadk_simt_kernel<LOG_P><<<(int)blocks, 256, smem, st>>>(a);
  return MQ_OK;
}

}  // namespace mq

namespace mq {
// shared by the two entry points: a.norm is the fp32 factor applied after the butterflies
static int hadamard_dispatch(const HadArgs& a, cudaStream_t st) {
  const int had_k = a.had_k;
  const int P = a.n / had_k;
  int logp = 0;
  while ((1 << logp) < P) ++logp;
  int rc = MQ_ERR_UNSUPPORTED;
  if (had_k == 1 && hadp2_tc_supported(a)) {
    rc = launch_hadp2_tc(a, st);  // H_A (x) H_64 on tcgen05, hi / lo split of the intermediate (mq_hadamard_p2.cu)
  } else if (had_k == 1 && fwht_mma_supported(a)) {
    rc = launch_fwht_mma(a, st);  // 16-point factors on the tensor pipe (mq_hadamard_mma.cu)
  } else if (had_k == 1) {
    switch (logp) {
      case 6: rc = launch_warp<6>(a, st); break;
      case 7: rc = launch_warp<7>(a, st); break;
      case 8: rc = launch_warp<8>(a, st); break;
      case 9: rc = launch_warp<9>(a, st); break;
      case 10: rc = launch_warp<10>(a, st); break;
      case 11: rc = launch_warp<11>(a, st); break;
      case 12: rc = launch_cta<12>(a, st); break;
      case 13: rc = launch_cta<13>(a, st); break;
      default:
        set_error("mq_hadamard: power-of-two n=%d outside [64, 8192]", a.n);
        return MQ_ERR_UNSUPPORTED;
    }
  } else if (hadk_tc3_supported(a) && (rc = launch_hadk_tc3(a, st)) != MQ_ERR_UNSUPPORTED) {
    // both Kronecker factors on tcgen05 (mq_hadamard_tc3.cu)
  } else if (hadk_tc2_supported(a) && (rc = launch_hadk_tc2(a, st)) != MQ_ERR_UNSUPPORTED) {
    // FWHT of the power-of-two factor in registers + hadK on tcgen05
  } else {
    switch (logp) {
      case 3: case 4: case 5:
        set_error("mq_hadamard_quant_i8: n/had_k=%d < 64 unsupported", P);
        return MQ_ERR_UNSUPPORTED;
      case 6: rc = launch_hadk<6>(a, st); break;
      case 7: rc = launch_hadk<7>(a, st); break;
      case 8: rc = launch_hadk<8>(a, st); break;
      case 9: rc = launch_hadk<9>(a, st); break;
      default:
        set_error("mq_hadamard_quant_i8: n/had_k=%d outside [64, 512]", P);
        return MQ_ERR_UNSUPPORTED;
    }
  }
  if (rc != MQ_OK) return rc;
  MQ_LAUNCH_CHECK("hadamard kernel");
  return MQ_OK;
}
}  // namespace mq

extern "C" int mq_hadamard_quant_i8(const void* x, int64_t ldx, int8_t* q, void* y_f16, float* srow_out,
                                    void* x0_out, const uint8_t* modality, const float* scales,
                                    const void* hadk, int had_k, int64_t M, int n_in, int n, unsigned flags,
                                    void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;  // empty batch
  MQ_REQUIRE(x != nullptr && (q != nullptr || y_f16 != nullptr), "mq_hadamard_quant_i8: null pointer");
  MQ_REQUIRE(q == nullptr || scales != nullptr, "mq_hadamard_quant_i8: codes requested without scales");
  MQ_REQUIRE(M >= 0 && n > 0 && n_in > 0 && n_in <= n && n_in % 8 == 0 && ldx % 8 == 0 && ldx >= n_in,
             "mq_hadamard_quant_i8: bad shape (n_in=%d n=%d ldx=%lld)", n_in, n, (long long)ldx);
  MQ_REQUIRE(had_k >= 1 && n % had_k == 0, "mq_hadamard_quant_i8: had_k=%d does not divide n=%d", had_k, n);
  MQ_REQUIRE(aligned16(x) && aligned16(q) && aligned16(y_f16), "mq_hadamard_quant_i8: pointers must be 16-byte aligned");
  MQ_REQUIRE(!(flags & MQ_FLAG_SPLIT) || (x0_out != nullptr && q != nullptr),
             "mq_hadamard_quant_i8: MQ_FLAG_SPLIT needs q and x0_out");
  const int P = n / had_k;
  MQ_REQUIRE((P & (P - 1)) == 0, "mq_hadamard_quant_i8: n/had_k=%d is not a power of two", P);
  MQ_REQUIRE(had_k == 1 || hadk != nullptr, "mq_hadamard_quant_i8: had_k>1 needs the hadK matrix");
  HadArgs a;
  a.x = (const __half*)x; a.ldx = ldx; a.q = q; a.y = (__half*)y_f16; a.srow_out = srow_out;
  a.x0_out = (__half*)x0_out; a.modality = modality; a.scales = scales; a.hadk = (const __half*)hadk;
  a.had_k = had_k; a.M = M; a.n_in = n_in; a.n = n; a.flags = flags;
  a.norm = 1.0f / sqrtf((float)n);  // float32(1/sqrt(n)): torch.tensor(n).sqrt() is fp32
  return hadamard_dispatch(a, (cudaStream_t)stream);
}

extern "C" int mq_hadamard_transform_f16(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, int n,
                                         float scale, void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;
  MQ_REQUIRE(x != nullptr && y != nullptr, "mq_hadamard_transform_f16: null pointer");
  MQ_REQUIRE(M >= 0 && n >= 64 && n <= 8192 && (n & (n - 1)) == 0,
             "mq_hadamard_transform_f16: n=%d must be a power of two in [64, 8192]", n);
  MQ_REQUIRE(ldx % 8 == 0 && ldx >= n && ldy == n, "mq_hadamard_transform_f16: bad strides (ldx=%lld ldy=%lld)",
             (long long)ldx, (long long)ldy);
  MQ_REQUIRE(aligned16(x) && aligned16(y), "mq_hadamard_transform_f16: pointers must be 16-byte aligned");
  HadArgs a;
  a.x = (const __half*)x; a.ldx = ldx; a.q = nullptr; a.y = (__half*)y; a.srow_out = nullptr;
  a.x0_out = nullptr; a.modality = nullptr; a.scales = nullptr; a.hadk = nullptr;
  a.had_k = 1; a.M = M; a.n_in = n; a.n = n; a.flags = 0;
  a.norm = scale;
  return hadamard_dispatch(a, (cudaStream_t)stream);
}
