// Activation quantisation kernels (HBM-bound, one pass).
//
//   mq_act_quant_i8          static INT8, MSQ two-scale select, zero-pad, RMS split
//   mq_act_fake_quant_f16    quant -> dequant in one pass (drop-in for BaseQuantizer.forward)
//   mq_act_quant_i8_dynamic  per-token symmetric scales + codes (ActQuantizer.find_params)
//
// Algorithmic bytes per element: 2 read + 1 written (codes) or 2 + 2 (fake quant).
// Each thread moves 16 halfs (two 128-bit loads) and writes one 128-bit vector of codes, so
// a warp reads 1 KB and writes 512 B fully coalesced; grid = a multiple of the SM count,
// grid-stride over (row, 16-column chunk).
#include "mq_common.cuh"

namespace mq {

// Work item = one 16-element chunk (two 128-bit loads, one 128-bit store).  Each thread takes
// kUnroll chunks per trip and issues all of their loads before any arithmetic, so 4 x 16 B per
// thread (~128 KB per SM at full occupancy) are in flight to cover HBM latency.
constexpr int kQuantUnroll = 2;

__global__ void __launch_bounds__(256)
act_quant_i8_kernel(const __half* __restrict__ x, int64_t ldx, int8_t* __restrict__ q,
                    float* __restrict__ srow_out, __half* __restrict__ x0_out,
                    const uint8_t* __restrict__ modality, const float* __restrict__ scales,
                    uint32_t total, FastDiv cpr, int n_in, int n_out, unsigned flags) {
  const QScale q_text = make_qscale(__ldg(scales)), q_vis = make_qscale(__ldg(scales + 1));
  // fp16 / bf16-valued scales (every MinmaxObserver scale): branch-free quantiser
  const bool narrow = q_text.narrow && q_vis.narrow;
  const QNarrow n_text = make_qnarrow(q_text), n_vis = make_qnarrow(q_vis);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += stride * kQuantUnroll) {
    uint4 va[kQuantUnroll], vb[kQuantUnroll];
    uint32_t row[kQuantUnroll];
    int col[kQuantUnroll];
    bool vis[kQuantUnroll];
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      const uint32_t c = c0 + u * stride;
      va[u] = vb[u] = make_uint4(0u, 0u, 0u, 0u);
      row[u] = 0;
      col[u] = 0;
      vis[u] = false;
      if (c < total) {
        row[u] = cpr.div(c);
        col[u] = (int)(c - row[u] * cpr.d) << 4;
        const __half* xr = x + (int64_t)row[u] * ldx + col[u];
        if (col[u] + 8 <= n_in) va[u] = ldg_stream(xr);
        if (col[u] + 16 <= n_in) vb[u] = ldg_stream(xr + 8);
        vis[u] = (modality != nullptr) && (__ldg(modality + row[u]) != 0);
      }
    }
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      const uint32_t c = c0 + u * stride;
      if (c >= total) break;
      const QScale& qs = vis[u] ? q_vis : q_text;
      uint4 out;
      uint2 r, r2;  // zero-padded columns quantise to code 0
      if (narrow) {
        const QNarrow& qn = vis[u] ? n_vis : n_text;
        r = quant8_narrow(va[u], qn);
        r2 = quant8_narrow(vb[u], qn);
      } else {
        r = quant8_bracket(va[u], qs);
        r2 = quant8_bracket(vb[u], qs);
      }
      out.x = r.x;
      out.y = r.y;
      out.z = r2.x;
      out.w = r2.y;
      if (col[u] == 0) {
        if (flags & MQ_FLAG_SPLIT) {
          x0_out[row[u]] = reinterpret_cast<const __half*>(&va[u])[0];
          out.x &= 0xffffff00u;
        }
        if (srow_out != nullptr) srow_out[row[u]] = qs.s;
      }
      stg_stream(q + (int64_t)row[u] * n_out + col[u], out);
    }
  }
}

// INT8 fake quantisation with fp16- / bf16-valued scales (every MinmaxObserver scale): the codes of the
// division-free quantiser (quant2_narrow, mq_common.cuh) never leave the fp32 domain -- the packed sum of
// the two bracket ends is 2 (M + code), so code = sum / 2 - M exactly -- and are dequantised with the
// reference's expression fp16(fp32(code) * s).  A zero result takes the sign of x + 0 (torch.round keeps
// the sign of a tiny negative quotient, "+ zero_point" turns -0 into +0).  ~6 instructions per element
// against ~14 of the generic kernel below; same results bit for bit.
__device__ __forceinline__ uint4 fake8_narrow(const uint4& v, const QNarrow& q, uint64_t s2) {
  uint4 out;
  uint32_t* o = reinterpret_cast<uint32_t*>(&out);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  uint64_t magic2, half2_, negm2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(magic2) : "f"(MQ_RND_MAGIC));
  asm("mov.b64 %0, {%1, %1};" : "=l"(half2_) : "f"(0.5f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(negm2) : "f"(-MQ_RND_MAGIC));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 xz = __hadd2(h[i], __float2half2_rn(0.f));  // -0 -> +0, everything else unchanged
    const float2 f = __half22float2(__hmin2(__hmax2(h[i], q.lim_lo), q.lim_hi));
    uint64_t x, lo, hi, sum, c, y;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(f.x), "f"(f.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(lo) : "l"(x), "l"(q.inv_lo2), "l"(magic2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(hi) : "l"(x), "l"(q.inv_hi2), "l"(magic2));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum) : "l"(lo), "l"(hi));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(sum), "l"(half2_), "l"(negm2));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(c), "l"(s2));
    float y0, y1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y));
    const __half2 r = __floats2half2_rn(y0, y1);
    o[i] = *reinterpret_cast<const uint32_t*>(&r) | (*reinterpret_cast<const uint32_t*>(&xz) & 0x80008000u);
  }
  return out;
}

__device__ __forceinline__ void fake_i8_narrow_body(const __half* __restrict__ x, int64_t ldx, __half* __restrict__ y,
                                                    int64_t ldy, const uint8_t* __restrict__ modality,
                                                    const QScale& q_text, const QScale& q_vis, uint32_t total,
                                                    const FastDiv& cpr) {
  const QNarrow n_text = make_qnarrow(q_text), n_vis = make_qnarrow(q_vis);
  uint64_t s2_text, s2_vis;
  asm("mov.b64 %0, {%1, %1};" : "=l"(s2_text) : "f"(q_text.s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(s2_vis) : "f"(q_vis.s));
  const uint32_t stride = gridDim.x * blockDim.x;
  // work item = one 16-byte chunk; four of them (64 B of loads) in flight per thread
  for (uint32_t c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += stride * 4) {
    uint4 v[4];
    uint32_t row[4];
    int col[4];
    bool vis[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t c = c0 + u * stride;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      row[u] = 0;
      col[u] = 0;
      vis[u] = false;
      if (c < total) {
        row[u] = cpr.div(c);
        col[u] = (int)(c - row[u] * cpr.d) << 3;
        v[u] = ldg_stream(x + (int64_t)row[u] * ldx + col[u]);
        vis[u] = (modality != nullptr) && (__ldg(modality + row[u]) != 0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c0 + u * stride >= total) break;
      stg_stream(y + (int64_t)row[u] * ldy + col[u], fake8_narrow(v[u], vis[u] ? n_vis : n_text, vis[u] ? s2_vis : s2_text));
    }
  }
}

// scale_mode 0: per-token scale by modality; 1: per-channel scale array
template <int kScaleMode>
__device__ __forceinline__ void fake_generic_body(const __half* __restrict__ x, int64_t ldx, __half* __restrict__ y,
                                                  int64_t ldy, const uint8_t* __restrict__ modality,
                                                  const float* __restrict__ scales, float lo, float hi, int64_t M, int C) {
  const int chunks_per_row = C >> 3;
  const int64_t total = M * chunks_per_row;
  float s_text = 0.f, s_vis = 0.f;
  if (kScaleMode == 0) {
    s_text = __ldg(scales);
    s_vis = __ldg(scales + 1);
  }
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < total;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = c / chunks_per_row;
    const int col = (int)(c - m * chunks_per_row) << 3;
    uint4 v = ldg_stream(x + m * ldx + col);
    float f[8];
    unpack8(v, f);
    if (kScaleMode == 0) {
      const float s = (modality != nullptr && modality[m] != 0) ? s_vis : s_text;
      const float inv = __frcp_rn(s);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __fmul_rn(quant_code(f[i], s, inv, lo, hi), s);
    } else {
      const float4 sa = __ldg(reinterpret_cast<const float4*>(scales + col));
      const float4 sb = __ldg(reinterpret_cast<const float4*>(scales + col + 4));
      const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __fmul_rn(quant_code(f[i], sc[i], __frcp_rn(sc[i]), lo, hi), sc[i]);
    }
    stg_stream(y + m * ldy + col, pack8_f16(f));
  }
}

template <int kScaleMode>
__global__ void __launch_bounds__(256)
act_fake_quant_kernel(const __half* __restrict__ x, int64_t ldx, __half* __restrict__ y, int64_t ldy,
                      const uint8_t* __restrict__ modality, const float* __restrict__ scales,
                      float lo, float hi, int64_t M, int C) {
  fake_generic_body<kScaleMode>(x, ldx, y, ldy, modality, scales, lo, hi, M, C);
}

// INT8, per-token scale by modality: the division-free body when both scales are fp16- / bf16-valued
// (they live in device memory, so the choice is made here, once per kernel), else the generic one
__global__ void __launch_bounds__(256)
act_fake_quant_i8_kernel(const __half* __restrict__ x, int64_t ldx, __half* __restrict__ y, int64_t ldy,
                         const uint8_t* __restrict__ modality, const float* __restrict__ scales, uint32_t total,
                         FastDiv cpr, int64_t M, int C) {
  const QScale q_text = make_qscale(__ldg(scales)), q_vis = make_qscale(__ldg(scales + 1));
  if (q_text.narrow && q_vis.narrow)
    fake_i8_narrow_body(x, ldx, y, ldy, modality, q_text, q_vis, total, cpr);
  else
    fake_generic_body<0>(x, ldx, y, ldy, modality, scales, -128.f, 127.f, M, C);
}

// One warp per token row: pass 1 min/max (row stays in L1/L2), pass 2 quantise.
__global__ void __launch_bounds__(256)
act_quant_i8_dynamic_kernel(const __half* __restrict__ x, int64_t ldx, int8_t* __restrict__ q,
                            float* __restrict__ srow_out, int64_t M, int n_in, int n_out,
                            float clip_ratio) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps) {
    const __half* xr = x + m * ldx;
    __half2 mx = __float2half2_rn(0.f), mn = __float2half2_rn(0.f);
    for (int col = lane << 3; col < n_in; col += 256) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(xr + col));
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mx = __hmax2(mx, h[i]);
        mn = __hmin2(mn, h[i]);
      }
    }
    float fmx = fmaxf(__low2float(mx), __high2float(mx));
    float fmn = fminf(__low2float(mn), __high2float(mn));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
      fmn = fminf(fmn, __shfl_xor_sync(0xffffffffu, fmn, o));
    }
    // quant_utils.py:240-247: xmin/xmax (with 0) * clip_ratio in fp32, xmax = max(|xmin|, xmax),
    // scale = xmax / maxq, all-zero rows -> 1
    const float xmin = __fmul_rn(fmn, clip_ratio), xmax0 = __fmul_rn(fmx, clip_ratio);
    const float xmax = fmaxf(fabsf(xmin), xmax0);
    const float s = (xmax == 0.f) ? 1.f : __fdiv_rn(xmax, 127.f);
    if (lane == 0) srow_out[m] = s;
    const QScale qs = make_qscale(s);
    for (int col = lane << 4; col < n_out; col += 512) {
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (col + 8 <= n_in) {
        uint2 r = quant8_bracket(__ldg(reinterpret_cast<const uint4*>(xr + col)), qs);
        out.x = r.x;
        out.y = r.y;
      }
      if (col + 16 <= n_in) {
        uint2 r = quant8_bracket(__ldg(reinterpret_cast<const uint4*>(xr + col + 8)), qs);
        out.z = r.x;
        out.w = r.y;
      }
      stg_stream(q + m * (int64_t)n_out + col, out);
    }
  }
}

// Same contract with the row held in REGISTERS between the two phases (one HBM read, no re-read through
// L1 / L2): WPT warps share a token, a thread owns up to 8 units of 16 consecutive elements (two 128-bit
// loads, all of them in flight together), min / max meet by shuffles and -- for WPT > 1 -- in shared
// memory.  NU = 4 units per thread wherever 8 warps can hold the row that way (~70 registers, three CTAs
// per SM; with 8 units and one warp per 4096-wide token the kernel was SLOWER than the two-pass one: 16
// warps per SM do not hide the quantiser behind the loads).  Rows up to 32768 elements; the kernel above
// serves anything wider.
template <int WPT, int NU>
__global__ void __launch_bounds__(256)
act_quant_i8_dynamic_reg_kernel(const __half* __restrict__ x, int64_t ldx, int8_t* __restrict__ q,
                                float* __restrict__ srow_out, int64_t M, int n_in, int n_out, float clip_ratio) {
  constexpr int TPC = 8 / WPT;  // tokens per CTA pass
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tslot = w / WPT, wsub = w % WPT;
  const int units = n_out >> 4;
  __shared__ float smx[8], smn[8];
  const int64_t passes = (M + TPC - 1) / TPC;
  for (int64_t pass = blockIdx.x; pass < passes; pass += gridDim.x) {
    const int64_t m = pass * TPC + tslot;
    const bool row_ok = m < M;
    const __half* xr = x + m * ldx;
    uint4 v[NU][2];
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      const int col = (wsub * 32 + lane + j * 32 * WPT) << 4;
      v[j][0] = (row_ok && col + 8 <= n_in) ? ldg_stream(xr + col) : make_uint4(0u, 0u, 0u, 0u);
      v[j][1] = (row_ok && col + 16 <= n_in) ? ldg_stream(xr + col + 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    __half2 mx = __float2half2_rn(0.f), mn = __float2half2_rn(0.f);
#pragma unroll
    for (int j = 0; j < NU; ++j) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[j][h2]);
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
    if (WPT > 1) {
      if (lane == 0) {
        smx[w] = fmx;
        smn[w] = fmn;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        fmx = fmaxf(fmx, smx[tslot * WPT + i]);
        fmn = fminf(fmn, smn[tslot * WPT + i]);
      }
      __syncthreads();  // (the slots are rewritten in the next pass)
    }
    // quant_utils.py:240-247: xmin/xmax (with 0) * clip_ratio in fp32, xmax = max(|xmin|, xmax),
    // scale = xmax / maxq, all-zero rows -> 1
    const float xmin = __fmul_rn(fmn, clip_ratio), xmax0 = __fmul_rn(fmx, clip_ratio);
    const float xmax = fmaxf(fabsf(xmin), xmax0);
    const float s = (xmax == 0.f) ? 1.f : __fdiv_rn(xmax, 127.f);
    if (row_ok && wsub == 0 && lane == 0) srow_out[m] = s;
    const QScale qs = make_qscale(s);
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      const int u = wsub * 32 + lane + j * 32 * WPT;
      if (row_ok && u < units) {
        const int col = u << 4;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (col + 8 <= n_in) {
          const uint2 r = quant8_bracket(v[j][0], qs);
          out.x = r.x;
          out.y = r.y;
        }
        if (col + 16 <= n_in) {
          const uint2 r = quant8_bracket(v[j][1], qs);
          out.z = r.x;
          out.w = r.y;
        }
        stg_stream(q + m * (int64_t)n_out + col, out);
      }
    }
  }
}

template <int WPT, int NU>
static void launch_dynamic_reg(const __half* x, int64_t ldx, int8_t* q, float* srow_out, int64_t M, int n_in, int n_out,
                               float clip_ratio, cudaStream_t st) {
  const int64_t passes = (M + (8 / WPT) - 1) / (8 / WPT);
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(passes < cap ? passes : cap);
  act_quant_i8_dynamic_reg_kernel<WPT, NU><<<grid, 256, 0, st>>>(x, ldx, q, srow_out, M, n_in, n_out, clip_ratio);
}

static int grid_for(int64_t work_items, int threads, int per_sm) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mq

extern "C" int mq_act_quant_i8(const void* x, int64_t ldx, int8_t* q, float* srow_out, void* x0_out,
                               const uint8_t* modality, const float* scales, int64_t M, int n_in,
                               int n_out, unsigned flags, void* stream) {
  if (M == 0) return MQ_OK;  // empty batch
  MQ_REQUIRE(x && q && scales, "mq_act_quant_i8: null pointer");
  MQ_REQUIRE(M >= 0 && n_in > 0 && n_out >= n_in, "mq_act_quant_i8: bad shape M=%lld n_in=%d n_out=%d",
             (long long)M, n_in, n_out);
  MQ_REQUIRE(n_in % 8 == 0 && n_out % 16 == 0 && ldx % 8 == 0 && ldx >= n_in,
             "mq_act_quant_i8: need n_in%%8==0, n_out%%16==0, ldx%%8==0 (n_in=%d n_out=%d ldx=%lld)", n_in,
             n_out, (long long)ldx);
  MQ_REQUIRE(mq::aligned16(x) && mq::aligned16(q), "mq_act_quant_i8: pointers must be 16-byte aligned");
  MQ_REQUIRE(!(flags & MQ_FLAG_SPLIT) || x0_out, "mq_act_quant_i8: MQ_FLAG_SPLIT needs x0_out");
  if (M == 0) return MQ_OK;
  const int64_t chunks = M * (n_out / 16);
  MQ_REQUIRE(chunks < (1ll << 31), "mq_act_quant_i8: M*n_out/16 = %lld exceeds 2^31 (split the batch)", (long long)chunks);
  const int grid = mq::grid_for((chunks + mq::kQuantUnroll - 1) / mq::kQuantUnroll, 256, 8);
  mq::act_quant_i8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, ldx, q, srow_out, (__half*)x0_out, modality, scales, (uint32_t)chunks,
      mq::FastDiv((uint32_t)(n_out / 16)), n_in, n_out, flags);
  MQ_LAUNCH_CHECK("act_quant_i8_kernel");
  return MQ_OK;
}

extern "C" int mq_act_fake_quant_f16(const void* x, int64_t ldx, void* y, int64_t ldy,
                                     const uint8_t* modality, const float* scales, int scale_mode,
                                     int bits, int64_t M, int C, void* stream) {
  if (M == 0) return MQ_OK;  // empty batch
  MQ_REQUIRE(x && y && scales, "mq_act_fake_quant_f16: null pointer");
  MQ_REQUIRE(M >= 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C,
             "mq_act_fake_quant_f16: need C%%8==0 and 8-element aligned strides (C=%d)", C);
  MQ_REQUIRE(mq::aligned16(x) && mq::aligned16(y), "mq_act_fake_quant_f16: pointers must be 16-byte aligned");
  MQ_REQUIRE(bits >= 2 && bits <= 24, "mq_act_fake_quant_f16: bits=%d out of range", bits);
  MQ_REQUIRE(scale_mode == 0 || scale_mode == 1, "mq_act_fake_quant_f16: scale_mode=%d", scale_mode);
  if (M == 0) return MQ_OK;
  const float lo = -(float)(1 << (bits - 1)), hi = (float)((1 << (bits - 1)) - 1);
  const int64_t chunks = M * (C / 8);
  const int grid = mq::grid_for(chunks, 256, 8);
  if (scale_mode == 0 && bits == 8 && chunks < (1ll << 31)) {
    mq::act_fake_quant_i8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __half*)x, ldx, (__half*)y, ldy, modality, scales, (uint32_t)chunks, mq::FastDiv((uint32_t)(C / 8)), M, C);
    MQ_LAUNCH_CHECK("act_fake_quant_i8_kernel");
    return MQ_OK;
  }
  if (scale_mode == 0)
    mq::act_fake_quant_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __half*)x, ldx, (__half*)y, ldy, modality, scales, lo, hi, M, C);
  else
    mq::act_fake_quant_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __half*)x, ldx, (__half*)y, ldy, modality, scales, lo, hi, M, C);
  MQ_LAUNCH_CHECK("act_fake_quant_kernel");
  return MQ_OK;
}

extern "C" int mq_act_quant_i8_dynamic(const void* x, int64_t ldx, int8_t* q, float* srow_out, int64_t M,
                                       int n_in, int n_out, float clip_ratio, void* stream) {
  if (M == 0) return MQ_OK;  // empty batch
  MQ_REQUIRE(x && q && srow_out, "mq_act_quant_i8_dynamic: null pointer");
  MQ_REQUIRE(M >= 0 && n_in > 0 && n_out >= n_in && n_in % 8 == 0 && n_out % 16 == 0 && ldx % 8 == 0 &&
                 ldx >= n_in,
             "mq_act_quant_i8_dynamic: bad shape/alignment (n_in=%d n_out=%d ldx=%lld)", n_in, n_out,
             (long long)ldx);
  MQ_REQUIRE(mq::aligned16(x) && mq::aligned16(q), "mq_act_quant_i8_dynamic: pointers must be 16-byte aligned");
  if (M == 0) return MQ_OK;
  const int units = n_out / 16;
  if (units <= 32 * 8 * 8) {
    // the row stays in registers: as few warps per token as hold it in 4 units per thread
    const __half* xh = (const __half*)x;
    cudaStream_t st = (cudaStream_t)stream;
    if (units <= 128) mq::launch_dynamic_reg<1, 4>(xh, ldx, q, srow_out, M, n_in, n_out, clip_ratio, st);
    else if (units <= 256) mq::launch_dynamic_reg<2, 4>(xh, ldx, q, srow_out, M, n_in, n_out, clip_ratio, st);
    else if (units <= 512) mq::launch_dynamic_reg<4, 4>(xh, ldx, q, srow_out, M, n_in, n_out, clip_ratio, st);
    else if (units <= 1024) mq::launch_dynamic_reg<8, 4>(xh, ldx, q, srow_out, M, n_in, n_out, clip_ratio, st);
    else mq::launch_dynamic_reg<8, 8>(xh, ldx, q, srow_out, M, n_in, n_out, clip_ratio, st);
    MQ_LAUNCH_CHECK("act_quant_i8_dynamic_reg_kernel");
    return MQ_OK;
  }
  const int grid = mq::grid_for(M * 32, 256, 8);
  mq::act_quant_i8_dynamic_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, ldx, q, srow_out,
                                                                          M, n_in, n_out, clip_ratio);
  MQ_LAUNCH_CHECK("act_quant_i8_dynamic_kernel");
  return MQ_OK;
}
