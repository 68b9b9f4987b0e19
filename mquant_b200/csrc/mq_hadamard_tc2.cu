// Online Hadamard (non-power-of-two sizes) + static INT8 quantisation, second generation:
// the power-of-two factor on CUDA cores, the hadK factor on tcgen05.
//
//   y[tok, k'*P + j'] = sum_k hadK[k',k] * fp16( norm * sum_j x[tok, k*P + j] * H_P[j,j'] )
//   (matmul_hadU_cuda, fake_quant/hadamard_utils.py:115-128: FHT over the last dim of the
//    [-1, K, n/K] view, fp16 rounding, hadK @ ., fp16 rounding), then UniformQuantizer.quant.
//
// Why not both factors on the tensor core (mq_hadamard_tc.cu): each MMA result has to come back
// through the TMEM read port (64 B/clk per SM).  Two drains of 4 B per element are 1/8 clk per
// element -- the same as the HBM time of 3 B per element at 1/148 of 6.5 TB/s -- and the 128-row
// MMA1 tiles padded had_k = 172 to 256 rows on top: that kernel stalled at 38 % of the HBM
// roofline with neither the tensor pipe nor HBM busy.  The FWHT of 64 points is 6 adds per
// element; in registers it needs no shared-memory or TMEM round trip at all:
//
//   TMA      32 segments (64 fp16 = 128 B each) per 4 KB box, SWIZZLE_128B, zero-filled past
//            n_in / past M
//   xform    thread = one segment: 8 x LDS.128 -> fp32 -> 64-point butterfly in registers
//            (packed FADD2 for five of the six stages) [-> log2(P/64) shuffle stages] -> * norm ->
//            fp16 (first rounding point) -> 8 x STS.128 = one 128-byte row of T, the MN-major A
//            operand of the MMA: T[(segment), k, j' % 64]
//   MMA      D[(tok,j'), k'] = T . hadK^T     M=128 (2 segments), N=Kpad, K=Kpad, fp32 in TMEM
//   epilogue D -> registers -> fp16 (second rounding point) -> fp16 store and/or exact INT8
//            quantiser with the per-token MSQ scale -> HBM
//
// Persistent, one CTA per SM, 512 threads: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM
// allocation), warps 2-7 transform, warps 8-15 epilogue.
//
// Unit of work = SEGS = max(2, P/64) segments of 64 j' for every k: two tokens for P = 64, one
// token otherwise; it owns one T buffer and SEGS/2 accumulators of Kpad columns.
// HBM traffic per token: 2*n_in read + n (codes) [+ 2*n fp16] written.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace hadtc2 {
using namespace mq::tc;
#ifdef MQ_TC2_SLEEP
#define mbar_wait mbar_wait_sleep
#endif

constexpr int NUM_THREADS = 512;  // 4 warps per scheduler -> 128 registers per thread (the transform keeps 64 fp32 live)
constexpr int XF_WARP0 = 2, NUM_XF_WARPS = 6;
constexpr int EPI_WARP0 = 8, NUM_EPI_WARPS = 8;
constexpr int MAX_XSTAGES = 16;  // X ring: as deep as shared memory allows (bytes in flight bound the loads)
constexpr int X_STAGE = 4096;
constexpr int TMEM_COLS = 512;
constexpr int D2_STRIDE = 256;  // TMEM columns per accumulator set (two sets)

struct Cfg {
  int had_k, Kpad, kb2;  // Kpad = round_up(had_k, 16); kb2 = ceil(Kpad / 64) k-blocks of hadK
  int cpu;               // chunks (TMA boxes of 32 segments) per unit
  int acc_cols;          // TMEM columns between the accumulators of a unit
  int tb;                // T buffers
  int xs;                // X ring depth
  int t_buf;             // bytes
  int b2_off, x_off, t_off, bar_off, total;
};

struct Params {
  HadArgs a;
  Cfg c;
  unsigned dbg;  // bottleneck probes (MQ_PROBES builds only)
};
#ifdef MQ_PROBES
#define MQ_HPROBE(prm, mask) (((prm).dbg & (mask)) != 0u)
#else
#define MQ_HPROBE(prm, mask) false
#endif

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// MQ_HAD_DEBUG bit 256: CTA 0 writes clock64() stamps into the (garbage anyway) code buffer,
// 8 slots per unit, tools/had_probe.py --timeline prints them
__device__ __forceinline__ void stamp(const Params& prm, int lu, int slot) {
  if (MQ_HPROBE(prm, 256u) && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && lu < 64)
    reinterpret_cast<long long*>(prm.a.q)[lu * 8 + slot] = clock64();
}

// NS = P / 64 sub-segments per (token, k) row
template <int NS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hadk_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params prm) {
  constexpr int P = 64 * NS;
  constexpr int SEGS = NS < 2 ? 2 : NS;  // segments of 64 j' per unit
  constexpr int TPU = SEGS / NS;         // tokens per unit
  constexpr int MT = SEGS / 2;           // 128-row MMA tiles per unit
  constexpr int KPC = 32 / SEGS;         // k rows per chunk
  const HadArgs& a = prm.a;
  const Cfg& c = prm.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sB2 = base + c.b2_off, sX = base + c.x_off, sT = base + c.t_off;
  const uint32_t bar0 = base + c.bar_off;
  // barrier map: x_full[16] x_empty[16] t_full[2] t_empty[2] d2_full[2] d2_empty[2]
  constexpr int XSTAGES = MAX_XSTAGES;  // (barrier slots; c.xs of them are in use)
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (XSTAGES + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * XSTAGES + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * XSTAGES + 2 + s); };
  auto d2_full = [&](int s) { return bar0 + 8u * (2 * XSTAGES + 4 + s); };
  auto d2_empty = [&](int s) { return bar0 + 8u * (2 * XSTAGES + 6 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + c.bar_off + 8 * (2 * XSTAGES + 8));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int64_t n_units = (a.M + TPU - 1) / TPU;
  const int my_units = blockIdx.x < n_units ? (int)((n_units - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  // ---- one-time setup: barriers, TMEM, hadK ------------------------------------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < c.xs; ++s) {
      mbar_init(x_full(s), 1);
      mbar_init(x_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), c.cpu);  // one arrive per chunk (by the warp that transformed it)
      mbar_init(t_empty(s), 1);
      mbar_init(d2_full(s), 1);
      mbar_init(d2_empty(s), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  // B2 = hadK as [k' rows][k] K-major SWIZZLE_128B, zero padded to Kpad x (kb2*64)
  for (int id = threadIdx.x; id < c.Kpad * c.kb2 * 8; id += NUM_THREADS) {
    const int row = id / (c.kb2 * 8), rem = id % (c.kb2 * 8), kb = rem / 8, ch = rem % 8;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = kb * 64 + ch * 8 + 2 * e;
      uint32_t lo = 0, hi = 0;
      if (row < c.had_k) {
        if (k0 < c.had_k) lo = __half_as_ushort(a.hadk[row * c.had_k + k0]);
        if (k0 + 1 < c.had_k) hi = __half_as_ushort(a.hadk[row * c.had_k + k0 + 1]);
      }
      w[e] = lo | (hi << 16);
    }
    const uint32_t dst = sB2 + kb * (c.Kpad * 128) + row * 128 + ((uint32_t)(ch ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: chunks of 32 segments =====================
    // Lane l owns chunk l (l + 32, ...) of every unit: the lanes wait for their ring slots
    // together (one barrier round trip per unit instead of one per 4 KB box), then the boxes are
    // issued lane by lane -- the TMA instruction takes warp-uniform operands.
    // (ring slot / phase of a lane's chunk advance by cpu per unit: no divisions in the loops --
    // five integer divisions per chunk were a third of the transform warps' time)
    int xs0 = lane % c.xs;
    uint32_t xph0 = (uint32_t)(lane / c.xs) & 1u;
    for (int lu = 0; lu < my_units; ++lu) {
      const int64_t u = (int64_t)blockIdx.x + (int64_t)lu * gridDim.x;
      int xs = xs0;
      uint32_t xph = xph0;
      for (int c0 = 0; c0 < c.cpu; c0 += 32) {
        const int ci = c0 + lane;
        const bool mine = ci < c.cpu;
        if (mine) mbar_wait(x_empty(xs), xph ^ 1u);
        __syncwarp();
        const int cnt = min(32, c.cpu - c0);
        for (int i = 0; i < cnt; ++i) {
          if (lane == i) {
            if (MQ_HPROBE(prm, 32u)) {  // probe: no loads
              mbar_arrive(x_full(xs));
            } else {
              mbar_arrive_expect_tx(x_full(xs), X_STAGE);
              if (NS == 1) {
                // [M][K_in][64] box {64, 16, 2}: rows t*16 + kl = (token u*2+t, k = ci*16+kl)
                tma_load_3d(sX + xs * X_STAGE, &tmap_x, 0, ci * KPC, (int)(u * 2), x_full(xs));
              } else {
                // [M][K_in][NS][64] box {64, NS, KPC, 1}: rows kl*NS + sub
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                    ::"r"(sX + xs * X_STAGE), "l"(&tmap_x), "r"(0), "r"(0), "r"(ci * KPC), "r"((int)u), "r"(x_full(xs))
                    : "memory");
              }
            }
          }
          __syncwarp();
        }
        xs += 32;  // next group of 32 chunks of this unit
        while (xs >= c.xs) { xs -= c.xs; xph ^= 1u; }
      }
      xs0 += c.cpu;  // same chunk of the next unit
      while (xs0 >= c.xs) { xs0 -= c.xs; xph0 ^= 1u; }
      stamp(prm, lu, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: D = T . hadK^T =====================
    // F16 x F16 -> F32, A MN-major, B K-major, N = Kpad, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(c.Kpad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int lu = 0; lu < my_units; ++lu) {
      const int tb = c.tb == 2 ? (lu & 1) : 0, d2 = lu & 1;
      const uint32_t tph = (uint32_t)(c.tb == 2 ? (lu >> 1) : lu) & 1u, d2ph = (uint32_t)(lu >> 1) & 1u;
      mbar_wait(t_full(tb), tph);
      stamp(prm, lu, 2);
      mbar_wait(d2_empty(d2), d2ph ^ 1u);
      stamp(prm, lu, 3);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t lbo = (uint32_t)c.Kpad * 128u;  // between the two 64-j' segments of a tile
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t tm = tmem_base + (uint32_t)(d2 * D2_STRIDE + mt * c.acc_cols);
          const uint32_t tbase = sT + tb * c.t_buf + (uint32_t)(2 * mt) * lbo;
          for (int k = 0; k < c.Kpad / 16; ++k) {
            if (MQ_HPROBE(prm, 8u)) break;  // probe: no MMA
            const uint32_t boff = (uint32_t)(k / 4) * (uint32_t)(c.Kpad * 128) + (uint32_t)(k % 4) * 32u;
            umma_f16(tm, make_sw128_desc(tbase + (uint32_t)k * 2048u, lbo, 1024),
                     make_sw128_desc(sB2 + boff, 0, 1024), idesc, (uint32_t)(k != 0));
          }
        }
        umma_commit(t_empty(tb));
        umma_commit(d2_full(d2));
      }
      __syncwarp();
      stamp(prm, lu, 4);
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: D -> fp16 -> fp16 store / INT8 codes -> HBM ===================
    // Thread = one j' (TMEM lane), 16 consecutive k' per chunk.  The 16 codes of a thread belong to
    // 16 different output rows (stride P), so groups of 4 lanes transpose their 4x4 byte blocks
    // with two shuffle+PRMT steps and every lane then stores 32-bit words (4 consecutive j' of one
    // k').  fp16 output: for a fixed k' the 32 lanes hold 32 consecutive j' = one 64-byte store.
    const int quad = warp & 3, grp = (warp - EPI_WARP0) >> 2;  // lane quadrant, column-chunk parity
    const int mrow = quad * 32 + lane;
    const int chunks = c.Kpad / 16;
    const int la = lane & 3;
    const uint32_t sel1 = (la & 1) ? 0x3715u : 0x6240u;  // 4x4 byte transpose, step xor-1
    const uint32_t sel2 = (la & 2) ? 0x3276u : 0x5410u;  // ... step xor-2
    const bool want_q = a.q != nullptr, want_y = a.y != nullptr;
    QScale q_text = {}, q_vis = {};
    if (want_q) {
      q_text = make_qscale(__ldg(a.scales));
      q_vis = make_qscale(__ldg(a.scales + 1));
    }
    const bool split = (a.flags & MQ_FLAG_SPLIT) != 0;
    // the token of this thread's rows does not depend on the MMA tile (P = 64: one tile, two
    // tokens; P >= 128: one token); its modality flag is fetched one unit ahead of its use
    auto token_of = [&](int64_t u) { return (NS == 1) ? (u * 2 + (mrow >> 6)) : u; };
    auto is_visual = [&](int64_t u) {
      const int64_t tok = token_of(u);
      return u < n_units && tok < a.M && a.modality != nullptr && a.modality[tok] != 0;
    };
    bool vis_next = is_visual(blockIdx.x);
    int lu = 0;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++lu) {
      const int d2 = lu & 1;
      const uint32_t d2ph = (uint32_t)(lu >> 1) & 1u;
      const bool vis = vis_next;
      vis_next = is_visual(u + gridDim.x);
      const int64_t tok = token_of(u);
      const bool live = tok < a.M;
      QScale qs;  // field-wise select keeps it in registers
      qs.s = vis ? q_vis.s : q_text.s;
      qs.inv_lo = vis ? q_vis.inv_lo : q_text.inv_lo;
      qs.inv_hi = vis ? q_vis.inv_hi : q_text.inv_hi;
      qs.lim_lo = vis ? q_vis.lim_lo : q_text.lim_lo;
      qs.lim_hi = vis ? q_vis.lim_hi : q_text.lim_hi;
      qs.narrow = vis ? q_vis.narrow : q_text.narrow;
      if (live && (mrow & 63) == 0 && (NS == 1 || mrow == 0) && grp == 0 && a.srow_out != nullptr) a.srow_out[tok] = qs.s;
      mbar_wait(d2_full(d2), d2ph);
      if (warp == EPI_WARP0) stamp(prm, lu, 5);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int seg = 2 * mt + (mrow >> 6), jl = mrow & 63;
        const int j = (NS == 1) ? jl : (seg * 64 + jl);
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(d2 * D2_STRIDE + mt * c.acc_cols);
        int8_t* obase = want_q ? a.q + tok * (int64_t)a.n + (j - la) : nullptr;  // word-aligned: 4 consecutive j'
        __half* ybase = want_y ? a.y + tok * (int64_t)a.n + j : nullptr;
        // Two 16-column chunks (this warp's ch and ch + 2) are converted together: four
        // independent quantiser streams give the scheduler the ILP that two epilogue warps per
        // SM sub-partition cannot (the epilogue is latency bound, not issue bound).
        // (NARROW: the division-free quantiser, chosen once per unit -- a test inside quant8 would cut
        // the independent elements of a thread into basic blocks of 8)
        auto process2 = [&](auto narrow_tag, const uint32_t (&v0)[16], int ch0, const uint32_t (&v1)[16], bool two) {
          constexpr bool NARROW = decltype(narrow_tag)::value;
          const QNarrow qn = make_qnarrow(qs);
          if (MQ_HPROBE(prm, 2u)) return;  // probe: TMEM loads only
          uint32_t w[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ch = ch0 + (g >> 1) * 2;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(g < 2 ? v0[(g & 1) * 8 + e] : v1[(g & 1) * 8 + e]);
            const uint4 h = pack8_f16(f);  // second rounding point (hadK @ . -> fp16)
            if (want_y && live && (g < 2 || two)) {
              const __half* hh = reinterpret_cast<const __half*>(&h);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int kp = ch * 16 + (g & 1) * 8 + e;
                if (kp < c.had_k) ybase[(int64_t)kp * P] = hh[e];
              }
            }
            uint2 r = make_uint2(0u, 0u);
            if (want_q) r = NARROW ? quant8_narrow(h, qn) : quant8_bracket(h, qs);
            if (split && ch == 0 && g == 0 && j == 0) {
              if (live) a.x0_out[tok] = reinterpret_cast<const __half*>(&h)[0];
              r.x &= 0xffffff00u;
            }
            w[2 * g] = r.x;
            w[2 * g + 1] = r.y;
          }
          if (want_q) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t t1 = __shfl_xor_sync(0xffffffffu, w[i], 1);
              uint32_t x1 = __byte_perm(w[i], t1, sel1);
              uint32_t t2 = __shfl_xor_sync(0xffffffffu, x1, 2);
              const uint32_t x2 = __byte_perm(x1, t2, sel2);  // codes of k' = 16ch + 4(i%4) + la, j' = 4*(lane/4) .. +3
              const int kp = (ch0 + (i >> 2) * 2) * 16 + 4 * (i & 3) + la;
              if (live && kp < c.had_k && (i < 4 || two)) *reinterpret_cast<uint32_t*>(obase + (int64_t)kp * P) = x2;
            }
          }
        };
        uint32_t va[16], vb[16];
        int ch = grp;
        if (MQ_HPROBE(prm, 64u)) ch = chunks;  // probe: barrier handshake only
        for (; ch < chunks; ch += 4) {
          const bool two = ch + 2 < chunks;
          tmem_ld16_nowait(taddr + ch * 16, va);
          if (two) tmem_ld16_nowait(taddr + (ch + 2) * 16, vb);
          tmem_wait_ld();
          if (qs.narrow)
            process2(std::true_type{}, va, ch, vb, two);
          else
            process2(std::false_type{}, va, ch, vb, two);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(d2));
      if (warp == EPI_WARP0) stamp(prm, lu, 6);
    }
  } else if (warp >= XF_WARP0) {
    // ===================== transform: one thread = one 64-element segment =========================
    const int xw = warp - XF_WARP0;  // chunks g = xw, xw + NUM_XF_WARPS, ... ; chunk g sits in X stage g % xs
    const int sgi = (NS == 1) ? (lane >> 4) : (lane & (SEGS - 1));  // segment of the unit
    const int kl = (NS == 1) ? (lane & 15) : (lane / SEGS);
    const uint64_t norm2 = pack2(a.norm, a.norm);
    const int total_chunks = my_units * c.cpu;
    int lu = xw / c.cpu, ci = xw % c.cpu, xs = xw % c.xs;  // (xs >= NUM_XF_WARPS)
    uint32_t xph = 0;
    for (int g = xw; g < total_chunks; g += NUM_XF_WARPS) {
      const int tb = c.tb == 2 ? (lu & 1) : 0;
      const uint32_t tph = (uint32_t)(c.tb == 2 ? (lu >> 1) : lu) & 1u;
      const uint32_t xrow = sX + xs * X_STAGE + lane * 128;
      mbar_wait(x_full(xs), xph);
      uint64_t v[32];  // v[i] = elements (2i, 2i+1)
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        uint4 rawv = make_uint4(0u, 0u, 0u, 0u);
        if (!MQ_HPROBE(prm, 128u))
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(rawv.x), "=r"(rawv.y), "=r"(rawv.z), "=r"(rawv.w)
                     : "r"(xrow + ((uint32_t)(cc ^ (lane & 7)) << 4)));
        const uint32_t wds[4] = {rawv.x, rawv.y, rawv.z, rawv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wds[e]));
          v[cc * 4 + e] = pack2(f.x, f.y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty(xs));  // every lane's loads have been consumed above
      if (!MQ_HPROBE(prm, 16u)) {
        // stage "xor 1" inside the packed pairs: (a, b) -> (a + b, a - b)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float lo = __uint_as_float((uint32_t)v[i]), hi = __uint_as_float((uint32_t)(v[i] >> 32));
          v[i] = pack2(__fadd_rn(lo, hi), __fsub_rn(lo, hi));
        }
        // stages xor 2 .. xor 32 on packed registers (index bit s-1 of the pair index)
#pragma unroll
        for (int s = 0; s < 5; ++s) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (!(i & (1 << s))) {
              const uint64_t x0 = v[i], x1 = v[i | (1 << s)];
              v[i] = add2(x0, x1);
              v[i | (1 << s)] = sub2(x0, x1);
            }
          }
        }
        // sub-segments of one row sit in adjacent lanes: remaining log2(NS) stages by shuffle
#pragma unroll
        for (int m = 1; m < NS; m <<= 1) {
          const bool upper = (lane & m) != 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint64_t o = __shfl_xor_sync(0xffffffffu, v[i], m);
            v[i] = upper ? sub2(o, v[i]) : add2(v[i], o);
          }
        }
      }
      // first rounding point: fp16(norm * .), then one 128-byte row of T (chunk cc at cc ^ (k & 7))
      const int k = ci * KPC + kl;
      const uint32_t trow = sT + tb * c.t_buf + (uint32_t)sgi * (uint32_t)(c.Kpad * 128) + (uint32_t)k * 128u;
      mbar_wait(t_empty(tb), tph ^ 1u);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint64_t sc = mul2(v[cc * 4 + e], norm2);
          const __half2 hh = __floats2half2_rn(__uint_as_float((uint32_t)sc), __uint_as_float((uint32_t)(sc >> 32)));
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const uint32_t dst = trow + ((uint32_t)(cc ^ (k & 7)) << 4);
        if (!MQ_HPROBE(prm, 4u))
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
      }
      if (!MQ_HPROBE(prm, 128u)) fence_async_smem();
      __syncwarp();  // one arrival per chunk
      if (lane == 0) mbar_arrive(t_full(tb));
      if (ci == 0) stamp(prm, lu, 1);
      if (ci == c.cpu - 1) stamp(prm, lu, 7);
      ci += NUM_XF_WARPS;
      while (ci >= c.cpu) { ci -= c.cpu; ++lu; }
      xs += NUM_XF_WARPS;
      if (xs >= c.xs) { xs -= c.xs; xph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int NS>
static int launch(const HadArgs& a, cudaStream_t st) {
  constexpr int P = 64 * NS;
  constexpr int SEGS = NS < 2 ? 2 : NS;
  constexpr int TPU = SEGS / NS;
  constexpr int KPC = 32 / SEGS;
  Params prm;
  prm.a = a;
  Cfg& c = prm.c;
  c.had_k = a.had_k;
  c.Kpad = round_up(a.had_k, 16);
  c.kb2 = (c.Kpad + 63) / 64;
  c.cpu = c.Kpad / KPC;
  c.acc_cols = round_up(c.Kpad, 32);
  c.t_buf = SEGS * c.Kpad * 128;
  const int b2 = c.kb2 * c.Kpad * 128;
  const int budget = 227 * 1024 - 1024 - 256;
  const int b2r = round_up(b2, 1024);
  c.tb = 2;
  if (b2r + NUM_XF_WARPS * X_STAGE + c.tb * c.t_buf > budget) c.tb = 1;
  if (b2r + NUM_XF_WARPS * X_STAGE + c.tb * c.t_buf > budget) return MQ_ERR_UNSUPPORTED;
  c.xs = (budget - b2r - c.tb * c.t_buf) / X_STAGE;
  if (c.xs > MAX_XSTAGES) c.xs = MAX_XSTAGES;
  c.b2_off = 0;
  c.x_off = b2r;
  c.t_off = c.x_off + c.xs * X_STAGE;
  c.bar_off = c.t_off + c.tb * c.t_buf;
  c.total = c.bar_off + 8 * (2 * MAX_XSTAGES + 8) + 16;
#ifdef MQ_PROBES
  const char* dbg_env = getenv("MQ_HAD_DEBUG");
  prm.dbg = dbg_env ? (unsigned)atoi(dbg_env) : 0u;
#else
  prm.dbg = 0u;
#endif

  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  // rows k >= n_in / P (zero padding) and tokens >= M are out of bounds -> zero filled by TMA
  CUtensorMap tmap;
  CUresult cr;
  const cuuint32_t estride[4] = {1, 1, 1, 1};
  if (NS == 1) {
    const cuuint64_t gdim[3] = {64, (cuuint64_t)(a.n_in / 64), (cuuint64_t)a.M};
    const cuuint64_t gstride[2] = {128, (cuuint64_t)a.ldx * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)KPC, 2};
    cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.x, gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t gdim[4] = {64, (cuuint64_t)NS, (cuuint64_t)(a.n_in / P), (cuuint64_t)a.M};
    const cuuint64_t gstride[3] = {128, (cuuint64_t)P * 2, (cuuint64_t)a.ldx * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)NS, (cuuint32_t)KPC, 1};
    cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)a.x, gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (cr != CUDA_SUCCESS) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  const int smem = c.total + 1024;
  MQ_DYN_SMEM(hadk_tc2_kernel<NS>, smem);
  const int64_t units = (a.M + TPU - 1) / TPU;
  const int grid = (int)(units < num_sms() ? units : num_sms());
  hadk_tc2_kernel<NS><<<grid, NUM_THREADS, smem, st>>>(tmap, prm);
  return MQ_OK;
}

}  // namespace hadtc2

// had_k > 1, fp16 rounding points of matmul_hadU_cuda: codes and/or fp16 output
bool hadk_tc2_supported(const HadArgs& a) {
  const int P = a.n / a.had_k;
  if (a.had_k <= 1 || (P != 64 && P != 128 && P != 256 && P != 512)) return false;
  if (a.q == nullptr && a.y == nullptr) return false;
  if (a.flags & MQ_FLAG_FP32_HAD) return false;
  if (a.n_in % P != 0 || a.M >= (1ll << 31)) return false;
  const int Kpad = (a.had_k + 15) / 16 * 16;
  const int mt = P / 64 < 2 ? 1 : P / 128;
  if (mt * ((Kpad + 31) / 32 * 32) > hadtc2::D2_STRIDE) return false;  // accumulators of a unit: 256 TMEM columns
  return forced_hadamard_kernel() != 1;
}

int launch_hadk_tc2(const HadArgs& a, cudaStream_t st) {
  switch (a.n / a.had_k) {
    case 64: return hadtc2::launch<1>(a, st);
    case 128: return hadtc2::launch<2>(a, st);
    case 256: return hadtc2::launch<4>(a, st);
    default: return hadtc2::launch<8>(a, st);
  }
}

}  // namespace mq
