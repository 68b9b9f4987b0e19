// Online Hadamard (non-power-of-two sizes) + static INT8 quantisation, entirely on tcgen05.
//
//   y[tok, k'*P + j'] = sum_k hadK[k',k] * fp16( norm * sum_j x[tok, k*P + j] * H_P[j,j'] )
//   (matmul_hadU_cuda, fake_quant/hadamard_utils.py:115-128: FHT over the last dim of the
//    [-1, K, n/K] view, fp16 rounding, hadK @ ., fp16 rounding), then UniformQuantizer.quant.
//
// Both Kronecker factors are +-1 matrices, so both run as fp16 MMAs with fp32 accumulation:
//
//   TMA      x rows (tok,k) [128 x P] fp16, SWIZZLE_128B, zero-filled past n_in / past M
//   MMA1     D1[(tok,k), j'] = X . H_P^T            M=128, N=P, K=P     (TMEM, fp32)
//   mid      D1 -> registers -> * float32(1/sqrt(n)) -> fp16 (the reference's first rounding
//            point) -> shared memory as the MN-major A operand of MMA2: T[(tok,j'), k]
//   MMA2     D2[(tok,j'), k'] = T . hadK^T          M=128 (2 x 64 j'), N=Kpad, K=Kpad
//   epilogue D2 -> registers -> fp16 (second rounding point) -> exact INT8 quantiser with the
//            per-token MSQ scale -> HBM
//
// Persistent, one CTA per SM, 512 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 mid stage, warps 8-15 epilogue.  The 172 x 172 factor alone is
// 3.8 MFLOP/token: on CUDA cores that kernel runs at 2.5 % of the HBM roofline (profiles/), here
// the tensor pipe needs ~1/2 of the per-token HBM time.
//
// HBM traffic per token: 2*n_in read + n written.  hadK / H_P live in shared memory for the whole
// kernel (written once per CTA in the K-major SWIZZLE_128B layout).
#include <cuda.h>

#include <cstdlib>

#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace hadtc {
using namespace mq::tc;

constexpr int NUM_THREADS = 512;
constexpr int MID_WARP0 = 4;
constexpr int EPI_WARP0 = 8;
constexpr int NUM_MID = 128;
constexpr int NUM_EPI = 256;
constexpr int TMEM_COLS = 512;
// TMEM plan (512 columns): two D2 accumulators (<= 176 columns each, at 0 and 192) + a 128-column
// D1 ring.  (Tried: ONE D2 accumulator + a 256-column D1 ring so that all MMA1s of a unit issue
// back to back -- 213 -> 257 us at M = 16384: MMA2 and the epilogue then serialise, and the
// probes showed the bare barrier/TMA chain (103 us) is bound by bytes in flight, not by the D1
// ping-pong.  profiles/r1_hadk_tc_probes.log)
constexpr int D2B = 2;
constexpr int D2_COL0 = 0, D1_COL = 384, D1_COLS = 128;

struct Cfg {
  int had_k, Kpad, kb2;   // Kpad = round_up(had_k, 16); kb2 = ceil(Kpad / 64) k-blocks of hadK
  int rt;                 // 128-row tiles per token in MMA1 = ceil(had_k / 128)
  int xs, tb;             // X ring depth, T buffers
  int b1_off, b2_off, x_off, t_off, bar_off, total;  // shared-memory carve-up (bytes)
  int x_stage, t_buf;     // bytes
};

struct Params {
  HadArgs a;
  Cfg c;
  unsigned dbg;
};

template <int P>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hadk_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const Params prm) {
  constexpr int S = P / 64;            // 64-wide j' segments per token
  constexpr int TPU = 2 / S;           // tokens per unit (one MMA2 tile = 2 segments)
  constexpr int D1B = D1_COLS / P;     // D1 ring depth
  constexpr int KB1 = P / 64;          // k-blocks of the MMA1 operands
  const HadArgs& a = prm.a;
  const Cfg& c = prm.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sB1 = base + c.b1_off, sB2 = base + c.b2_off, sX = base + c.x_off, sT = base + c.t_off;
  const uint32_t bar0 = base + c.bar_off;
  // barrier map: x_full[xs] x_empty[xs] d1_full[4] d1_empty[4] t_full[2] t_empty[2] d2_full[2] d2_empty[2]
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (c.xs + s); };
  auto d1_full = [&](int s) { return bar0 + 8u * (2 * c.xs + s); };
  auto d1_empty = [&](int s) { return bar0 + 8u * (2 * c.xs + 4 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * c.xs + 8 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * c.xs + 10 + s); };
  auto d2_full = [&](int s) { return bar0 + 8u * (2 * c.xs + 12 + s); };
  auto d2_empty = [&](int s) { return bar0 + 8u * (2 * c.xs + 14 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + c.bar_off + 8 * (2 * c.xs + 16));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int rtu = c.rt * TPU;  // row tiles per unit
  const int64_t n_units = (a.M + TPU - 1) / TPU;

  // ---- one-time setup: barriers, TMEM, the two constant operands --------------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < c.xs; ++s) {
      mbar_init(x_full(s), 1);
      mbar_init(x_empty(s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(d1_full(s), 1);
      mbar_init(d1_empty(s), NUM_MID / 32);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), (NUM_MID / 32) * rtu);
      mbar_init(t_empty(s), 1);
      mbar_init(d2_full(s), 1);
      mbar_init(d2_empty(s), NUM_EPI / 32);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  // B1 = H_P as [j' rows][j] K-major: 16-byte chunk (row, kb, ch) holds j = kb*64 + ch*8 .. +7
  for (int id = threadIdx.x; id < P * KB1 * 8; id += NUM_THREADS) {
    const int row = id / (KB1 * 8), rem = id % (KB1 * 8), kb = rem / 8, ch = rem % 8;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j0 = kb * 64 + ch * 8 + 2 * e;
      const uint32_t lo = (__popc(j0 & row) & 1) ? 0xBC00u : 0x3C00u;        // -1.0h / +1.0h
      const uint32_t hi = (__popc((j0 + 1) & row) & 1) ? 0xBC00u : 0x3C00u;
      w[e] = lo | (hi << 16);
    }
    const uint32_t dst = sB1 + kb * (P * 128) + row * 128 + ((uint32_t)(ch ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
  // B2 = hadK as [k' rows][k] K-major, zero padded to Kpad x (kb2*64)
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

  // instruction descriptors (cute::UMMA::InstrDescriptor): F16 x F16 -> F32
  const uint32_t idesc1 = (1u << 4) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (1u << 15) /* A is MN-major */ | ((uint32_t)(c.Kpad >> 3) << 17) |
                          ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    // ===================== TMA producer: X row tiles =====================
    if (lane == 0) {
      int xs = 0;
      uint32_t xph = 0;
      for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        for (int tk = 0; tk < TPU; ++tk) {
          const int tok = (int)(u * TPU + tk);
          for (int r = 0; r < c.rt; ++r) {
            mbar_wait(x_empty(xs), xph ^ 1u);
            mbar_arrive_expect_tx(x_full(xs), (uint32_t)c.x_stage);
#pragma unroll
            for (int kb = 0; kb < KB1; ++kb)
              tma_load_3d(sX + xs * c.x_stage + kb * (128 * 128), &tmap_x, kb * 64, r * 128, tok, x_full(xs));
            if (++xs == c.xs) { xs = 0; xph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int xs = 0, d1 = 0;
    uint32_t xph = 0, d1ph = 0;
    auto issue_mma1 = [&](int count) {  // `count` consecutive row tiles
      for (int i = 0; i < count; ++i) {
        mbar_wait(x_full(xs), xph);
        mbar_wait(d1_empty(d1), d1ph ^ 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t tm = tmem_base + D1_COL + d1 * P;
#pragma unroll
          for (int k = 0; k < P / 16; ++k) {
            if (prm.dbg & 16u) break;  // probe: no MMA1
            const uint32_t koff = (uint32_t)(k / 4) * (128 * 128) + (uint32_t)(k % 4) * 32u;
            const uint32_t boff = (uint32_t)(k / 4) * (P * 128) + (uint32_t)(k % 4) * 32u;
            umma_f16(tm, make_sw128_desc(sX + xs * c.x_stage + koff, 0, 1024),
                     make_sw128_desc(sB1 + boff, 0, 1024), idesc1, (uint32_t)(k != 0));
          }
          umma_commit(x_empty(xs));
          umma_commit(d1_full(d1));
        }
        __syncwarp();
        if (++xs == c.xs) { xs = 0; xph ^= 1u; }
        if (++d1 == D1B) { d1 = 0; d1ph ^= 1u; }
      }
    };
    const int my_units = (int)((n_units - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int pre = rtu < D1B ? rtu : D1B;  // row tiles of the NEXT unit issued ahead of MMA2
    if (my_units > 0) issue_mma1(rtu);
    for (int lu = 0; lu < my_units; ++lu) {
      const int tb = lu % c.tb, d2 = lu % D2B;
      const uint32_t tph = (uint32_t)(lu / c.tb) & 1u, d2ph = (uint32_t)(lu / D2B) & 1u;
      const bool more = lu + 1 < my_units;
      if (more) issue_mma1(pre);
      mbar_wait(t_full(tb), tph);
      mbar_wait(d2_empty(d2), d2ph ^ 1u);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t tm = tmem_base + D2_COL0 + d2 * 192;
        const uint32_t tbase = sT + tb * c.t_buf;
        const uint32_t lbo = (prm.dbg & 1u) ? 1024u : (uint32_t)c.Kpad * 128u;
        const uint32_t sbo = (prm.dbg & 1u) ? (uint32_t)c.Kpad * 128u : 1024u;
        for (int k = 0; k < c.Kpad / 16; ++k) {
          if (prm.dbg & 8u) break;  // probe: no MMA2
          const uint32_t boff = (uint32_t)(k / 4) * (uint32_t)(c.Kpad * 128) + (uint32_t)(k % 4) * 32u;
          umma_f16(tm, make_sw128_desc(tbase + (uint32_t)k * 2048u, lbo, sbo),
                   make_sw128_desc(sB2 + boff, 0, 1024), idesc2, (uint32_t)(k != 0));
        }
        umma_commit(t_empty(tb));
        umma_commit(d2_full(d2));
      }
      __syncwarp();
      if (more) issue_mma1(rtu - pre);
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue: D2 -> fp16 -> INT8 codes -> HBM =====================
    // Thread = one j' (TMEM lane), 16 consecutive k' per chunk.  The 16 codes of a thread belong to
    // 16 different output rows (stride P), so groups of 4 lanes transpose their 4x4 byte blocks
    // with two shuffle+PRMT steps and every lane then stores 32-bit words (4 consecutive j' of one
    // k'): 4 STG.32 per chunk instead of 16 byte stores.  TMEM loads are double buffered.
    const int quad = warp & 3, grp = (warp - EPI_WARP0) >> 2;  // lane quadrant, column-chunk parity
    const int mrow = quad * 32 + lane;                         // D2 lane = (segment, j' within segment)
    const int seg = mrow >> 6, jl = mrow & 63;
    const int chunks = c.Kpad / 16;
    const int la = lane & 3;
    const uint32_t sel1 = (la & 1) ? 0x3715u : 0x6240u;  // 4x4 byte transpose, step xor-1
    const uint32_t sel2 = (la & 2) ? 0x3276u : 0x5410u;  // ... step xor-2
    const QScale q_text = make_qscale(__ldg(a.scales)), q_vis = make_qscale(__ldg(a.scales + 1));
    const bool split = (a.flags & MQ_FLAG_SPLIT) != 0;
    int lu = 0;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++lu) {
      const int d2 = lu % D2B;
      const uint32_t d2ph = (uint32_t)(lu / D2B) & 1u;
      const int64_t tok = (S == 1) ? (u * TPU + seg) : u;
      const int j = (S == 1) ? jl : (seg * 64 + jl);
      const bool live = tok < a.M;
      const bool vis = live && a.modality != nullptr && a.modality[tok] != 0;
      QScale qs;  // field-wise select keeps it in registers
      qs.s = vis ? q_vis.s : q_text.s;
      qs.inv_lo = vis ? q_vis.inv_lo : q_text.inv_lo;
      qs.inv_hi = vis ? q_vis.inv_hi : q_text.inv_hi;
      qs.lim_lo = vis ? q_vis.lim_lo : q_text.lim_lo;
      qs.lim_hi = vis ? q_vis.lim_hi : q_text.lim_hi;
      qs.narrow = vis ? q_vis.narrow : q_text.narrow;
      if (live && j == 0 && grp == 0 && a.srow_out != nullptr) a.srow_out[tok] = qs.s;
      mbar_wait(d2_full(d2), d2ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D2_COL0 + d2 * 192);
      int8_t* obase = a.q + tok * (int64_t)a.n + (j - la);  // word-aligned: 4 consecutive j' per store
      auto process = [&](const uint32_t (&v)[16], int ch) {
        if (prm.dbg & 2u) return;  // probe: TMEM loads only
        uint32_t w[4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]);
          const uint4 h = pack8_f16(f);  // second rounding point (hadK @ . -> fp16)
          uint2 r = quant8(h, qs);
          if (split && ch == 0 && g == 0 && j == 0) {
            if (live) a.x0_out[tok] = reinterpret_cast<const __half*>(&h)[0];
            r.x &= 0xffffff00u;
          }
          w[2 * g] = r.x;
          w[2 * g + 1] = r.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t t1 = __shfl_xor_sync(0xffffffffu, w[i], 1);
          uint32_t x1 = __byte_perm(w[i], t1, sel1);
          uint32_t t2 = __shfl_xor_sync(0xffffffffu, x1, 2);
          const uint32_t x2 = __byte_perm(x1, t2, sel2);  // codes of k' = 16ch + 4i + la, j' = 4*(lane/4) .. +3
          const int kp = ch * 16 + 4 * i + la;
          if (live && kp < c.had_k) *reinterpret_cast<uint32_t*>(obase + (int64_t)kp * P) = x2;
        }
      };
      uint32_t va[16], vb[16];
      int ch = grp;
      if (ch < chunks) tmem_ld16_nowait(taddr + ch * 16, va);
      for (; ch < chunks; ch += 4) {
        tmem_wait_ld();
        if (ch + 2 < chunks) tmem_ld16_nowait(taddr + (ch + 2) * 16, vb);
        process(va, ch);
        if (ch + 2 < chunks) {
          tmem_wait_ld();
          if (ch + 4 < chunks) tmem_ld16_nowait(taddr + (ch + 4) * 16, va);
          process(vb, ch + 2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(d2));
    }
  } else if (warp >= MID_WARP0) {
    // ===================== mid stage: D1 -> * norm -> fp16 -> T (MN-major A of MMA2) =============
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // row of the 128-row tile
    int d1 = 0, lu = 0;
    uint32_t d1ph = 0;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++lu) {
      const int tb = lu % c.tb;
      const uint32_t tph = (uint32_t)(lu / c.tb) & 1u;
      mbar_wait(t_empty(tb), tph ^ 1u);
      for (int tk = 0; tk < TPU; ++tk) {
        for (int r = 0; r < c.rt; ++r) {
          const int k = r * 128 + row;
          mbar_wait(d1_full(d1), d1ph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D1_COL + d1 * P);
          auto convert = [&](const uint32_t (&v)[32], int cb) {
            if (k < c.Kpad && !(prm.dbg & 4u)) {
              // segment of these 32 j' values; T[(seg), k, j' % 64]
              const int sgi = tk * S + (cb * 32) / 64;
              const uint32_t trow = sT + tb * c.t_buf + (uint32_t)sgi * (uint32_t)(c.Kpad * 128) + (uint32_t)k * 128u;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __fmul_rn(__uint_as_float(v[g * 8 + e]), a.norm);
                const uint4 h = pack8_f16(f);  // first rounding point (FHT output -> fp16)
                const int chn = ((cb * 32) % 64) / 8 + g;
                const uint32_t dst = trow + ((uint32_t)(chn ^ (k & 7)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w) : "memory");
              }
            }
          };
          // TMEM loads double buffered against the conversion
          uint32_t va[32], vb[32];
          tmem_ld32_nowait(taddr, va);
#pragma unroll
          for (int cb = 0; cb < P / 32; cb += 2) {
            tmem_wait_ld();
            tmem_ld32_nowait(taddr + (cb + 1) * 32, vb);
            convert(va, cb);
            tmem_wait_ld();
            if (cb + 2 < P / 32) tmem_ld32_nowait(taddr + (cb + 2) * 32, va);
            convert(vb, cb + 1);
          }
          tc_fence_before();
          fence_async_smem();
          __syncwarp();  // one arrival per warp on each barrier
          if (lane == 0) {
            mbar_arrive(d1_empty(d1));
            mbar_arrive(t_full(tb));
          }
          if (++d1 == D1B) { d1 = 0; d1ph ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int P>
static int launch(const HadArgs& a, cudaStream_t st) {
  Params prm;
  prm.a = a;
  Cfg& c = prm.c;
  c.had_k = a.had_k;
  c.Kpad = round_up(a.had_k, 16);
  c.kb2 = (c.Kpad + 63) / 64;
  c.rt = (a.had_k + 127) / 128;
  c.x_stage = 128 * P * 2;
  c.t_buf = 2 * c.Kpad * 128;
  const int b1 = P * P * 2, b2 = c.kb2 * c.Kpad * 128;
  const int budget = 227 * 1024 - 1024 - 256;
  // prefer two T buffers and a 3-deep X ring; shrink until it fits
  c.tb = 2;
  c.xs = 3;
  while (b1 + b2 + c.xs * c.x_stage + c.tb * c.t_buf > budget) {
    if (c.xs > 2) --c.xs;
    else if (c.tb > 1) --c.tb;
    else {
      set_error("mq_hadamard_quant_i8: had_k=%d P=%d does not fit in shared memory", a.had_k, P);
      return MQ_ERR_UNSUPPORTED;
    }
  }
  c.b1_off = 0;
  c.b2_off = c.b1_off + b1;
  c.x_off = c.b2_off + b2;
  c.t_off = c.x_off + c.xs * c.x_stage;
  c.bar_off = c.t_off + c.tb * c.t_buf;
  c.total = c.bar_off + 8 * (2 * c.xs + 16) + 16;
  const char* dbg_env = getenv("MQ_HAD_DEBUG");
  prm.dbg = dbg_env ? (unsigned)atoi(dbg_env) : 0u;

  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled is not available from the driver");
    return MQ_ERR_CUDA;
  }
  // x viewed as [M tokens][n_in / P rows][P] fp16; rows >= n_in / P (zero padding) and tokens >= M
  // are out of bounds -> zero filled by the TMA unit
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)P, (cuuint64_t)(a.n_in / P), (cuuint64_t)a.M};
  const cuuint64_t gstride[2] = {(cuuint64_t)P * 2, (cuuint64_t)a.ldx * 2};
  const cuuint32_t box[3] = {64, 128, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.x, gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("mq_hadamard_quant_i8: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return MQ_ERR_CUDA;
  }
  const int smem = c.total + 1024;
  MQ_CUDA(cudaFuncSetAttribute(hadk_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  constexpr int TPU = 2 / (P / 64);
  const int64_t units = (a.M + TPU - 1) / TPU;
  const int grid = (int)(units < num_sms() ? units : num_sms());
  hadk_tc_kernel<P><<<grid, NUM_THREADS, smem, st>>>(tmap, prm);
  return MQ_OK;
}

}  // namespace hadtc

// had_k > 1, codes only: returns MQ_ERR_UNSUPPORTED (without setting an error) when the shape is
// outside this kernel so that the caller can fall back to the CUDA-core baseline.
bool hadk_tc_supported(const HadArgs& a) {
  const int P = a.n / a.had_k;
  if (a.had_k <= 1 || (P != 64 && P != 128)) return false;
  if (a.q == nullptr || a.y != nullptr) return false;
  if (a.flags & MQ_FLAG_FP32_HAD) return false;
  if (a.n_in % P != 0 || a.M >= (1ll << 31)) return false;
  const char* off = getenv("MQ_HAD_SIMT");
  if (off && atoi(off)) return false;
  return true;
}

int launch_hadk_tc(const HadArgs& a, cudaStream_t st) {
  const int P = a.n / a.had_k;
  return P == 64 ? hadtc::launch<64>(a, st) : hadtc::launch<128>(a, st);
}

}  // namespace mq
