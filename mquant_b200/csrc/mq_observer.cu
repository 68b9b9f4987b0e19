// Calibration observers (HBM-bound reductions over fp16 activations).
//
//   mq_observer_minmax       MinmaxObserver.update (fake_quant/observer/minmax.py:13-28)
//   mq_observer_percentile   the two torch.quantile calls of PercentileObserver.update
//                            (fake_quant/observer/percentile.py:31-32), exact
//
// minmax: algorithmic bytes = 2*M*C read, O(1) / O(C) written.  max/min of fp16 values are exact
// and order independent, so the running statistics can be merged with atomics across CTAs,
// streams and (after an NCCL allreduce-max) GPUs and stay bit-identical to the sequential
// reference.  Running max >= 0 and running min <= 0 always hold (the reference folds 0 in on
// the first update), which lets both be merged with integer atomics on the float bit pattern:
//   max: non-negative floats order like signed ints      -> atomicMax(int)
//   min: non-positive floats order like unsigned ints    -> atomicMax(unsigned) (more negative
//        = larger pattern; +0.0 = 0 is the identity)
#include "mq_common.cuh"

namespace mq {

__device__ __forceinline__ void merge_max(float* addr, float v) {
  if (v > 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
__device__ __forceinline__ void merge_min(float* addr, float v) {
  if (v < 0.f) atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// layer_wise: one scalar pair for the whole [M, C] tensor
__global__ void __launch_bounds__(256)
minmax_tensor_kernel(const __half* __restrict__ x, int64_t ldx, int64_t M, int C, float* __restrict__ io) {
  const int chunks_per_row = C >> 3;
  const int64_t total = M * chunks_per_row;
  __half2 mx = __float2half2_rn(0.f), mn = __float2half2_rn(0.f);
  const bool contiguous = (ldx == C);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // 4 independent 128-bit loads in flight per thread (pure streaming reduction: latency bound
  // otherwise)
  for (int64_t c0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c0 < total; c0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t c = c0 + u * stride;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (c < total) {
        const __half* p;
        if (contiguous) {
          p = x + c * 8;
        } else {
          const int64_t m = c / chunks_per_row;
          p = x + m * ldx + ((c - m * chunks_per_row) << 3);
        }
        v[u] = ldg_stream(p);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mx = __hmax2(mx, h[i]);
        mn = __hmin2(mn, h[i]);
      }
    }
  }
  float fmx = fmaxf(__low2float(mx), __high2float(mx));
  float fmn = fminf(__low2float(mn), __high2float(mn));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
    fmn = fminf(fmn, __shfl_xor_sync(0xffffffffu, fmn, o));
  }
  __shared__ float smx[8], smn[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    smx[w] = fmx;
    smn[w] = fmn;
  }
  __syncthreads();
  if (w == 0) {
    fmx = lane < 8 ? smx[lane] : 0.f;
    fmn = lane < 8 ? smn[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
      fmn = fminf(fmn, __shfl_xor_sync(0xffffffffu, fmn, o));
    }
    if (lane == 0) {
      merge_min(io + 0, fmn);
      merge_max(io + 1, fmx);
    }
  }
}

// layer_wise with Modality-Specific Static scales: one pass over x updates the text pair and the
// visual pair (row m belongs to modality[m]) and counts the rows of either kind -- the host never
// has to gather x[mask] / x[~mask] or to ask the device whether a modality occurred.
// Thread = (row, 8-element chunk); rows are walked by warps so that the modality byte is a
// warp-uniform broadcast load.
__global__ void __launch_bounds__(256)
minmax_tensor_msq_kernel(const __half* __restrict__ x, int64_t ldx, int64_t M, int C,
                         const uint8_t* __restrict__ modality, float* __restrict__ io_text,
                         float* __restrict__ io_vis, unsigned* __restrict__ seen) {
  const int chunks_per_row = C >> 3;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp_id = (int64_t)blockIdx.x * 8 + w, n_warps = (int64_t)gridDim.x * 8;
  __half2 mx[2] = {__float2half2_rn(0.f), __float2half2_rn(0.f)};
  __half2 mn[2] = {__float2half2_rn(0.f), __float2half2_rn(0.f)};
  unsigned rows[2] = {0u, 0u};
  for (int64_t m = warp_id; m < M; m += n_warps) {
    const int kind = modality[m] != 0 ? 1 : 0;
    rows[kind] += (lane == 0);
    const __half* xr = x + m * ldx;
    __half2 rmx = __float2half2_rn(0.f), rmn = rmx;
    // 4 independent 128-bit loads in flight per thread
    for (int c0 = lane; c0 < chunks_per_row; c0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        v[u] = c < chunks_per_row ? ldg_stream(xr + ((int64_t)c << 3)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rmx = __hmax2(rmx, h[i]);
          rmn = __hmin2(rmn, h[i]);
        }
      }
    }
    mx[kind] = __hmax2(mx[kind], rmx);
    mn[kind] = __hmin2(mn[kind], rmn);
  }
  __shared__ float smx[2][8], smn[2][8];
  __shared__ unsigned srows[2][8];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float fmx = fmaxf(__low2float(mx[k]), __high2float(mx[k]));
    float fmn = fminf(__low2float(mn[k]), __high2float(mn[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
      fmn = fminf(fmn, __shfl_xor_sync(0xffffffffu, fmn, o));
    }
    if (lane == 0) {
      smx[k][w] = fmx;
      smn[k][w] = fmn;
      srows[k][w] = rows[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int k = threadIdx.x;
    float fmx = 0.f, fmn = 0.f;
    unsigned r = 0;
    for (int i = 0; i < 8; ++i) {
      fmx = fmaxf(fmx, smx[k][i]);
      fmn = fminf(fmn, smn[k][i]);
      r += srows[k][i];
    }
    float* io = k ? io_vis : io_text;
    merge_min(io + 0, fmn);
    merge_max(io + 1, fmx);
    if (r) atomicAdd(seen + k, r);
  }
}

// channel_wise: io = {min[C], max[C]}.  blockIdx.x -> 256-channel slab (32 lanes x 8 channels: a warp
// reads 512 contiguous bytes of a row), the 8 warps of a CTA take different rows (blockIdx.y -> row
// slice), 4 independent 128-bit loads in flight per thread; the warps' partial results meet in shared
// memory and the CTA issues ONE pair of atomics per channel.  (The first version kept 2048 channels per
// CTA and one row in flight per thread, and every thread merged its 8 channels with 16 global atomics:
// 4.8 M atomics on 8192 words for 16384 x 4096 -- 0.25 of HBM.)
__global__ void __launch_bounds__(256)
minmax_channel_kernel(const __half* __restrict__ x, int64_t ldx, int64_t M, int C, float* __restrict__ io) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + lane) << 3;
  const bool col_ok = col < C;
  __half2 mx[4], mn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) mx[i] = mn[i] = __float2half2_rn(0.f);
  const int64_t row0 = (int64_t)blockIdx.y * 8 + w, step = (int64_t)gridDim.y * 8;
  if (col_ok) {
    for (int64_t m = row0; m < M; m += 4 * step) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = m + u * step;
        v[u] = r < M ? ldg_stream(x + r * ldx + col) : make_uint4(0u, 0u, 0u, 0u);  // (0 is the neutral start value)
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mx[i] = __hmax2(mx[i], h[i]);
          mn[i] = __hmin2(mn[i], h[i]);
        }
      }
    }
  }
  __shared__ __half2 smx[8][32][4], smn[8][32][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    smx[w][lane][i] = mx[i];
    smn[w][lane][i] = mn[i];
  }
  __syncthreads();
  // thread t finishes channel t of the slab: column t / 8 (lane), element t % 8
  const int t = threadIdx.x, ln = t >> 3, e = t & 7;
  const int ch = blockIdx.x * 256 + t;
  if (ch < C) {
    float fmx = 0.f, fmn = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      const __half2 a = smx[ww][ln][e >> 1], b = smn[ww][ln][e >> 1];
      fmx = fmaxf(fmx, (e & 1) ? __high2float(a) : __low2float(a));
      fmn = fminf(fmn, (e & 1) ? __high2float(b) : __low2float(b));
    }
    merge_min(io + ch, fmn);
    merge_max(io + C + ch, fmx);
  }
}

// ---------------------------------------------------------------------------------------------
// Percentile: exact order statistics of fp16 data by counting.
//   key(x)  = monotone 16-bit map of the fp16 pattern (negative: ~bits, non-negative: bits|0x8000)
//   pass 1  : 65536-bin histogram of keys.  Per-CTA privatisation does not fit (256 KB), so each
//             CTA histograms the TOP 12 bits in shared memory (4096 bins, warp-aggregated) and the
//             fine 16-bit histogram is built only for the coarse bins that contain a wanted rank
//             (pass 2 re-reads the tensor: 4*M*C bytes total, as SURVEY.md 8(d) budgets).
//   finish  : ranks floor/ceil(q*(n-1)) -> values -> torch's lerp, in fp32 (ATen quantile).
// workspace layout (uint32): coarse[4096] | fine[4][16] | sel[4] (selected coarse bins)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned key16(unsigned short b) {
  return (b & 0x8000u) ? (unsigned)(~b & 0xffffu) : (unsigned)(b | 0x8000u);
}
__device__ __forceinline__ float key_to_float(unsigned k) {
  const unsigned short b = (k & 0x8000u) ? (unsigned short)(k & 0x7fffu) : (unsigned short)(~k & 0xffffu);
  return __half2float(__ushort_as_half(b));
}

constexpr int kCoarseBins = 4096;
constexpr int kPctWords = kCoarseBins + 4 * 16 + 8;

__global__ void __launch_bounds__(256)
pct_coarse_kernel(const __half* __restrict__ x, int64_t ldx, int64_t M, int C, unsigned* __restrict__ ws) {
  __shared__ unsigned hist[kCoarseBins];
  for (int i = threadIdx.x; i < kCoarseBins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int chunks_per_row = C >> 3;
  const int64_t total = M * chunks_per_row;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < total;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = c / chunks_per_row;
    const uint4 v = ldg_stream(x + m * ldx + ((c - m * chunks_per_row) << 3));
    const unsigned short* h = reinterpret_cast<const unsigned short*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&hist[key16(h[i]) >> 4], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kCoarseBins; i += blockDim.x)
    if (hist[i]) atomicAdd(&ws[i], hist[i]);
}

// ranks wanted: r[0..3] = floor/ceil of the low rank, floor/ceil of the high rank
__device__ __forceinline__ void wanted_ranks(int64_t n, double alpha, int64_t (&r)[4], float (&w)[2]) {
  // torch.quantile: q as fp32 scalar, rank = q * (n - 1) in fp32
  const float q_hi = (float)alpha;
  const float q_lo = (float)(1.0 - alpha);
  const float nm1 = (float)(n - 1);
  const float rk_lo = __fmul_rn(q_lo, nm1), rk_hi = __fmul_rn(q_hi, nm1);
  r[0] = (int64_t)floorf(rk_lo);
  r[1] = (int64_t)ceilf(rk_lo);
  r[2] = (int64_t)floorf(rk_hi);
  r[3] = (int64_t)ceilf(rk_hi);
  w[0] = rk_lo - floorf(rk_lo);
  w[1] = rk_hi - floorf(rk_hi);
}

// One block of 1024 threads, 4 coarse bins each: block-wide exclusive scan of the counts, then every
// thread looks for the four wanted ranks inside its own bins.  (The first version walked the 4096
// bins with one thread -- 4096 dependent global loads, 0.3 ms of the 0.37 ms a small call took.)
__global__ void __launch_bounds__(1024) pct_select_kernel(unsigned* __restrict__ ws, int64_t n, double alpha) {
  static_assert(kCoarseBins == 4096, "4 bins per thread of a 1024-thread block");
  __shared__ unsigned long long warp_tot[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint4 c4 = reinterpret_cast<const uint4*>(ws)[t];
  const unsigned c[4] = {c4.x, c4.y, c4.z, c4.w};
  unsigned long long s = (unsigned long long)c[0] + c[1] + c[2] + c[3];
  unsigned long long inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned long long v = warp_tot[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += u;
    }
    warp_tot[lane] = iv - v;  // exclusive prefix of the warp totals
  }
  __syncthreads();
  int64_t cum = (int64_t)(warp_tot[w] + inc - s);  // elements in the bins before this thread's
  int64_t r[4];
  float wgt[2];
  wanted_ranks(n, alpha, r, wgt);
  unsigned* sel = ws + kCoarseBins + 64;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t next = cum + c[i];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (r[f] >= cum && r[f] < next) {
        sel[f] = (unsigned)(4 * t + i);
        sel[4 + f] = (unsigned)(r[f] - cum);  // rank of the wanted element inside this coarse bin
      }
    }
    cum = next;
  }
}

__global__ void __launch_bounds__(256)
pct_fine_kernel(const __half* __restrict__ x, int64_t ldx, int64_t M, int C, unsigned* __restrict__ ws) {
  __shared__ unsigned fine[4][16];
  if (threadIdx.x < 64) fine[threadIdx.x >> 4][threadIdx.x & 15] = 0;
  const unsigned* sel = ws + kCoarseBins + 64;
  const unsigned s0 = sel[0], s1 = sel[1], s2 = sel[2], s3 = sel[3];
  __syncthreads();
  const int chunks_per_row = C >> 3;
  const int64_t total = M * chunks_per_row;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < total;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = c / chunks_per_row;
    const uint4 v = ldg_stream(x + m * ldx + ((c - m * chunks_per_row) << 3));
    const unsigned short* h = reinterpret_cast<const unsigned short*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned k = key16(h[i]);
      const unsigned cb = k >> 4;
      // the four selected bins are the extreme tails: almost every element skips all four
      if (cb == s0) atomicAdd(&fine[0][k & 15], 1u);
      if (cb == s1) atomicAdd(&fine[1][k & 15], 1u);
      if (cb == s2) atomicAdd(&fine[2][k & 15], 1u);
      if (cb == s3) atomicAdd(&fine[3][k & 15], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const unsigned v = fine[threadIdx.x >> 4][threadIdx.x & 15];
    if (v) atomicAdd(&ws[kCoarseBins + threadIdx.x], v);
  }
}

__global__ void pct_finish_kernel(const unsigned* __restrict__ ws, int64_t n, double alpha,
                                  float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t r[4];
  float w[2];
  wanted_ranks(n, alpha, r, w);
  const unsigned* sel = ws + kCoarseBins + 64;
  float val[4];
  for (int i = 0; i < 4; ++i) {
    const unsigned* f = ws + kCoarseBins + 16 * i;
    unsigned want = sel[4 + i], cum = 0;
    unsigned k = 15;
    for (unsigned b = 0; b < 16; ++b) {
      cum += f[b];
      if (want < cum) {
        k = b;
        break;
      }
    }
    val[i] = key_to_float((sel[i] << 4) | k);
  }
  // at::lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w), each op one fp32 rounding
  for (int j = 0; j < 2; ++j) {
    const float lo = val[2 * j], hi = val[2 * j + 1];
    const float diff = __fsub_rn(hi, lo);
    out[j] = (w[j] < 0.5f) ? __fadd_rn(lo, __fmul_rn(w[j], diff))
                           : __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.f, w[j])));
  }
}

}  // namespace mq

extern "C" int mq_observer_minmax(const void* x, int64_t ldx, int64_t M, int C, int per_channel,
                                  float* minmax_io, void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;  // empty batch: statistics unchanged
  MQ_REQUIRE(x && minmax_io, "mq_observer_minmax: null pointer");
  MQ_REQUIRE(M >= 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C,
             "mq_observer_minmax: need C%%8==0 and ldx%%8==0 (C=%d ldx=%lld)", C, (long long)ldx);
  MQ_REQUIRE(aligned16(x), "mq_observer_minmax: x must be 16-byte aligned");
  if (M == 0) return MQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!per_channel) {
    const int64_t chunks = M * (C / 8);
    int64_t blocks = (chunks + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    minmax_tensor_kernel<<<(int)blocks, 256, 0, st>>>((const __half*)x, ldx, M, C, minmax_io);
  } else {
    const int slabs = (C + 255) / 256;
    int64_t slices = ((int64_t)num_sms() * 8 + slabs - 1) / slabs;  // ~8 CTAs per SM in total
    if (slices > (M + 7) / 8) slices = (M + 7) / 8;
    if (slices > 65535) slices = 65535;
    if (slices < 1) slices = 1;
    dim3 grid(slabs, (unsigned)slices);
    minmax_channel_kernel<<<grid, 256, 0, st>>>((const __half*)x, ldx, M, C, minmax_io);
  }
  MQ_LAUNCH_CHECK("minmax kernel");
  return MQ_OK;
}

extern "C" int mq_observer_minmax_msq(const void* x, int64_t ldx, int64_t M, int C, const uint8_t* modality,
                                      float* minmax_text_io, float* minmax_visual_io, unsigned* seen_io,
                                      void* stream) {
  using namespace mq;
  if (M == 0) return MQ_OK;  // empty batch: statistics unchanged
  MQ_REQUIRE(x && modality && minmax_text_io && minmax_visual_io && seen_io, "mq_observer_minmax_msq: null pointer");
  MQ_REQUIRE(M >= 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C,
             "mq_observer_minmax_msq: need C%%8==0 and ldx%%8==0 (C=%d ldx=%lld)", C, (long long)ldx);
  MQ_REQUIRE(aligned16(x), "mq_observer_minmax_msq: x must be 16-byte aligned");
  if (M == 0) return MQ_OK;
  int64_t blocks = (M + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  minmax_tensor_msq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, ldx, M, C, modality, minmax_text_io, minmax_visual_io, seen_io);
  MQ_LAUNCH_CHECK("minmax_tensor_msq_kernel");
  return MQ_OK;
}

extern "C" size_t mq_observer_percentile_workspace(void) { return sizeof(unsigned) * mq::kPctWords; }

extern "C" int mq_observer_percentile(const void* x, int64_t ldx, int64_t M, int C, double alpha, float* out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mq;
  MQ_REQUIRE(x && out && workspace, "mq_observer_percentile: null pointer");
  MQ_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C,
             "mq_observer_percentile: need M>0, C%%8==0, ldx%%8==0 (C=%d)", C);
  MQ_REQUIRE(aligned16(x), "mq_observer_percentile: x must be 16-byte aligned");
  MQ_REQUIRE(alpha > 0.5 && alpha < 1.0, "mq_observer_percentile: alpha=%f outside (0.5, 1)", alpha);
  if (workspace_bytes < mq_observer_percentile_workspace()) {
    set_error("mq_observer_percentile: workspace %zu < %zu bytes", workspace_bytes,
              mq_observer_percentile_workspace());
    return MQ_ERR_WORKSPACE;
  }
  // torch.quantile forms the rank in fp32: beyond 2^24 elements the reference itself leaves this
  // branch (percentile.py:33-43 numpy fallback); we keep the same arithmetic at any size.
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* ws = (unsigned*)workspace;
  MQ_CUDA(cudaMemsetAsync(ws, 0, mq_observer_percentile_workspace(), st));
  const int64_t n = M * (int64_t)C;
  const int64_t chunks = M * (C / 8);
  int64_t blocks = (chunks + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  pct_coarse_kernel<<<(int)blocks, 256, 0, st>>>((const __half*)x, ldx, M, C, ws);
  MQ_LAUNCH_CHECK("pct_coarse_kernel");
  pct_select_kernel<<<1, 1024, 0, st>>>(ws, n, alpha);
  MQ_LAUNCH_CHECK("pct_select_kernel");
  pct_fine_kernel<<<(int)blocks, 256, 0, st>>>((const __half*)x, ldx, M, C, ws);
  MQ_LAUNCH_CHECK("pct_fine_kernel");
  pct_finish_kernel<<<1, 32, 0, st>>>(ws, n, alpha, out);
  MQ_LAUNCH_CHECK("pct_finish_kernel");
  return MQ_OK;
}
