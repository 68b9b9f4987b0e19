// Online Walsh-Hadamard rotation for power-of-two sizes, fused with the static INT8 quantiser:
// the butterflies as 16-point matrix products on the tensor pipe (mma.sync), one warp per token.
//
//   y = x . H_n / sqrt(n)   (matmul_hadU_cuda with K == 1, fake_quant/hadamard_utils.py:117-120:
//   fast_hadamard_transform in fp32, one rounding to fp16), then UniformQuantizer.quant.
//
// Why: 12 butterfly stages on CUDA cores cost ~38 instructions per element (5 of the stages are
// shuffles on the LSU pipe) -- the kernel ran at 28 % of the HBM roofline, issue bound.  H_n is
// the Kronecker product of H_2 over the index bits, in ANY grouping of the bits, so a token of
// n = 256 * NB elements is transformed as
//
//   stage 1  4 index bits   D1 = H16 . X      mma.m16n8k16, X (exact fp16) is the B operand: one
//                                             128-bit shared-memory load IS the B fragment
//   stage 2  4 index bits   D2 = D1 . H16     D1 moves from the accumulator layout to the A layout
//                                             without data movement; to keep fp32 accuracy it is
//                                             split into fp16 hi + bf16 lo (2 x 2 mma, ~2^-19)
//   stage 3  log2(NB) bits  fp32 butterflies across the NB blocks, all in the registers of a thread
//
// and the accumulator layout leaves every thread with 4 consecutive outputs per tile: one
// 32-bit store of codes (or one 64-bit store of fp16).  ~13 instructions per element in total.
// Power-of-two parts of 1/sqrt(n) are folded into the +-1 matrices (exact), the remaining
// factor (sqrt(1/2) for odd log2 n) is one fp32 multiply, as in the reference.
//
// Index bits of an element inside its 256-block (b0..b7), lane = 4 g + t:
//   load   lane l reads the 16-byte chunk l:  b0,b1,b2 in the thread, (b3,b4) = t, (b5,b6,b7) = g
//   stage 1 contracts {b0, b2, b3, b4} (tile = b1), stage 2 contracts {b1, b5, b6, b7}
//   output thread holds (b0, b1, b7), g = (b2,b3,b4), t = (b5,b6): word g + 8 t + 32 tile of the block
//
// Tokens reach shared memory by cp.async.bulk (one per-warp double buffer, no CTA-wide barriers);
// the zero padding from n_in to n is a tail of the buffer that the copies never touch.
// HBM traffic per token: 2*n_in read + n written (codes) [+ 2*n if y_f16 is requested].
#include "mq_common.cuh"
#include "mq_hadamard_args.cuh"
#include "mq_tc.cuh"

namespace mq {
namespace hadmma {
using namespace mq::tc;

constexpr int WARPS = 4;

__device__ __forceinline__ void hmma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// fp32 pair -> fp16 pair (hi) and the bf16 pair of what the fp16 rounding dropped (lo, negated
// sign handled by the caller: this returns hi - v, one mixed-precision add per element)
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& neg_lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  float r0, r1;
  asm("{.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tsub.rn.f32.f16 %0, l, %3;\n\tsub.rn.f32.f16 %1, h, %4;}"
      : "=f"(r0), "=f"(r1)
      : "r"(hi), "f"(v0), "f"(v1));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(neg_lo) : "f"(r1), "f"(r0));
}

__device__ __forceinline__ uint32_t h2_bits(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t bf2_bits(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// per-token quantiser constants (narrow-scale fast path of mq_common.cuh + the scale itself)
struct QPack {
  QNarrow n;
  float s;
  bool narrow;
};
__device__ __forceinline__ QPack make_qpack(float s) {
  const QScale q = make_qscale(s);
  QPack p;
  p.n = make_qnarrow(q);
  p.s = s;
  p.narrow = q.narrow;
  return p;
}
// slow path: redo the 4 codes of a group with the true division
__device__ __noinline__ uint32_t quant4_exact(uint32_t a, uint32_t b, float s) {
  const __half2 ha = *reinterpret_cast<const __half2*>(&a), hb = *reinterpret_cast<const __half2*>(&b);
  return pack4_lowbytes(q8_exact(__low2float(ha), s), q8_exact(__high2float(ha), s), q8_exact(__low2float(hb), s),
                        q8_exact(__high2float(hb), s));
}

template <int NB, int FB_>
struct Traits {
  static constexpr int N = 256 * NB;
  static constexpr int LOG_N = NB == 4 ? 10 : NB == 8 ? 11 : NB == 16 ? 12 : 13;
  static constexpr int FOLD = LOG_N / 2;  // 2^-FOLD goes into the +-1 matrices (exact); what is left of
                                          // the scale (sqrt(1/2) for odd log2 n) is one fp32 multiply
  // Block-index bits: FB of them are folded into the stage-1 accumulation (CH = 2^FB input
  // blocks are summed with signs into every output group), the other log2(BL) are in-thread
  // butterflies over the BL blocks of a group.  Only BL * 8 accumulators are live at a time.
  static constexpr int FB = FB_;
  static constexpr int CH = 1 << FB;
  static constexpr int BL = NB / CH;
  static constexpr int MIN_CTAS = BL >= 16 ? 3 : BL >= 8 ? 4 : 5;
  // two token buffers per warp while 4 warps x MIN_CTAS x 2 x N halfs fit in shared memory: the
  // next token loads while this one is computed (measured: 8 us of exposed load latency at n = 4096)
  static constexpr int NBUF = (MIN_CTAS * 4 * 2 * N * 2 <= 200 * 1024) ? 2 : 1;
  static_assert(FOLD >= 4 + FB, "scale folding");
};

// MODE: bit 0 = codes, bit 1 = fp16 output.  NARROW: both scales take the division-free quantiser
// (decided once per kernel: a test per group of 4 codes cuts the epilogue into basic blocks of 4
// elements and serialises it)
template <int NB, int FB, int MODE, bool NARROW>
__device__ __forceinline__ void fwht_mma_body(const HadArgs& a) {
  using T = Traits<NB, FB>;
  constexpr int N = T::N, CH = T::CH, BL = T::BL, NBUF = T::NBUF;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // per warp: NBUF token buffers of N halfs, then (after all warps' buffers) NBUF mbarriers
  uint8_t* wbuf = smem + (size_t)warp * (NBUF * N * 2);
  const uint32_t buf0 = smem_u32(wbuf);
  const uint32_t bar0 = smem_u32(smem + (size_t)WARPS * (NBUF * N * 2)) + warp * 16;

  // ---- constant fragments ------------------------------------------------------------------------
  // A1[m][kk] = +-2^-(4+FB), sign = parity(m3 kk0 + m1 kk1 + m2 kk2 + m0 kk3)
  // B2[k2][n2] = +-2^-(FOLD-4-FB), sign = parity(k2_3 n2_0 + k2_0 n2_1 + k2_1 n2_2 + k2_2 n2_3)
  uint32_t a1p[4], b2h[2][2], b2b[2][2];
  {
    const float s1 = 1.0f / (float)(1 << (4 + T::FB));
    auto sgn1 = [s1](int m, int kk) {
      const int mp = ((m >> 3) & 1) | (m & 6) | ((m & 1) << 3);
      return (__popc(mp & kk) & 1) ? -s1 : s1;
    };
    const float s2 = 1.0f / (float)(1 << (T::FOLD - 4 - T::FB));
    auto sgn2 = [s2](int k2, int n2) {
      const int kp = ((k2 >> 3) & 1) | ((k2 & 7) << 1);
      return (__popc(kp & n2) & 1) ? -s2 : s2;
    };
    a1p[0] = h2_bits(sgn1(g, 2 * t), sgn1(g, 2 * t + 1));
    a1p[1] = h2_bits(sgn1(g + 8, 2 * t), sgn1(g + 8, 2 * t + 1));
    a1p[2] = h2_bits(sgn1(g, 2 * t + 8), sgn1(g, 2 * t + 9));
    a1p[3] = h2_bits(sgn1(g + 8, 2 * t + 8), sgn1(g + 8, 2 * t + 9));
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
      const int n2 = g + 8 * tl;
      // the lo operand arrives negated (split_pair returns hi - v): negate its matrix instead
      b2h[tl][0] = h2_bits(sgn2(2 * t, n2), sgn2(2 * t + 1, n2));
      b2h[tl][1] = h2_bits(sgn2(2 * t + 8, n2), sgn2(2 * t + 9, n2));
      b2b[tl][0] = bf2_bits(-sgn2(2 * t, n2), -sgn2(2 * t + 1, n2));
      b2b[tl][1] = bf2_bits(-sgn2(2 * t + 8, n2), -sgn2(2 * t + 9, n2));
    }
  }

  // ---- buffer: zero the padding tail once, barrier -----------------------------------------------
  const uint32_t row_bytes = (uint32_t)a.n_in * 2u;
  for (int b = 0; b < NBUF; ++b)
    for (int c = a.n_in / 8 + lane; c < N / 8; c += 32)
      *reinterpret_cast<uint4*>(wbuf + (size_t)b * N * 2 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t nw = (int64_t)gridDim.x * WARPS;
  QPack q_text = {}, q_vis = {};
  constexpr bool want_q = (MODE & 1) != 0, want_y = (MODE & 2) != 0;
  if (want_q) {
    q_text = make_qpack(__ldg(a.scales));
    q_vis = make_qpack(__ldg(a.scales + 1));
  }
  const float rest = a.norm * (float)(1 << T::FOLD);  // exact: a power-of-two multiple of the scale
  const uint64_t rest2 = f2_pack(rest, rest);
  const bool mul_rest = rest != 1.0f;
  const bool split = (a.flags & MQ_FLAG_SPLIT) != 0;

  if (gw < a.M && lane == 0) {
    mbar_arrive_expect_tx(bar0, row_bytes);
    bulk_load(buf0, a.x + gw * a.ldx, row_bytes, bar0);
  }
  const int wofs = g + 8 * t;  // word (4 elements) inside a block, tile adds 32
  int it = 0;
  for (int64_t m = gw; m < a.M; m += nw, ++it) {
    const int cur = NBUF == 2 ? (it & 1) : 0;
    const uint32_t ph = (uint32_t)(NBUF == 2 ? (it >> 1) : it) & 1u;
    const uint32_t src = buf0 + cur * (N * 2) + lane * 16;
    if (NBUF == 2 && m + nw < a.M && lane == 0) {
      // the other buffer was last read one iteration ago (every lane passed the __syncwarp since)
      mbar_arrive_expect_tx(bar0 + 8 * (cur ^ 1), row_bytes);
      bulk_load(buf0 + (cur ^ 1) * (N * 2), a.x + (m + nw) * a.ldx, row_bytes, bar0 + 8 * (cur ^ 1));
    }
    bool vis = false;
    if (want_q && a.modality != nullptr) vis = __ldg(a.modality + m) != 0;
    const QPack& qs = vis ? q_vis : q_text;
    if (lane == 0 && a.srow_out != nullptr && want_q) a.srow_out[m] = qs.s;
    int8_t* qrow = want_q ? a.q + m * (int64_t)a.n + 4 * wofs : nullptr;
    __half* yrow = want_y ? a.y + m * (int64_t)a.n + 4 * wofs : nullptr;
    mbar_wait(bar0 + 8 * cur, ph);

#pragma unroll 1
    for (int ch = 0; ch < CH; ++ch) {  // output group: blocks ch * BL .. ch * BL + BL - 1
      // (rolled: the unrolled body of all groups is ~60 KB of code and thrashes the instruction
      // caches once 20+ warps run out of phase.)  Signed stage-1 matrices of this group:
      // H_CH[ch][bh] * A1 for bh = 1 .. CH-1 (bh = 0 is always +)
      uint32_t a1s[CH > 1 ? CH - 1 : 1][4];
#pragma unroll
      for (int bh = 1; bh < CH; ++bh) {
        const uint32_t flip = (__popc(ch & bh) & 1) ? 0x80008000u : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) a1s[bh - 1][i] = a1p[i] ^ flip;
      }
      float acc[BL][2][4];
#pragma unroll
      for (int bl = 0; bl < BL; ++bl) {
        // stage 1 (+ the folded block bits): D1 = sum_bh (+-H16) . X[bh * BL + bl]
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int bh = 0; bh < CH; ++bh) {
          uint4 w;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                       : "r"(src + (bh * BL + bl) * 512));
          // tile 0 = words (0, 2), tile 1 = words (1, 3)
          hmma_f16(d0, bh == 0 ? a1p : a1s[bh > 0 ? bh - 1 : 0], w.x, w.z);
          hmma_f16(d1, bh == 0 ? a1p : a1s[bh > 0 ? bh - 1 : 0], w.y, w.w);
        }
        if (ch == CH - 1 && bl == BL - 1) {
          // last read of the buffer (single-buffer variant: refill it with this warp's next token
          // while the rest of this one is computed)
          __syncwarp();
          if (NBUF == 1 && m + nw < a.M && lane == 0) {
            mbar_arrive_expect_tx(bar0, row_bytes);
            bulk_load(buf0, a.x + (m + nw) * a.ldx, row_bytes, bar0);
          }
        }
        // stage 2: D1 as the A operand, fp16 hi + bf16 lo
        uint32_t ah[4], al[4];
        split_pair(d0[0], d0[1], ah[0], al[0]);
        split_pair(d0[2], d0[3], ah[1], al[1]);
        split_pair(d1[0], d1[1], ah[2], al[2]);
        split_pair(d1[2], d1[3], ah[3], al[3]);
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[bl][tl][e] = 0.f;
          hmma_bf16(acc[bl][tl], al, b2b[tl][0], b2b[tl][1]);
          hmma_f16(acc[bl][tl], ah, b2h[tl][0], b2h[tl][1]);
        }
      }

      // stage 3: butterflies across the blocks of the group, packed fp32 adds
#pragma unroll
      for (int s = 1; s < BL; s <<= 1) {
#pragma unroll
        for (int b = 0; b < BL; ++b) {
          if (b & s) continue;
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              const uint64_t x0 = f2_pack(acc[b][tl][2 * p], acc[b][tl][2 * p + 1]);
              const uint64_t x1 = f2_pack(acc[b + s][tl][2 * p], acc[b + s][tl][2 * p + 1]);
              uint64_t su, di;
              asm("add.rn.f32x2 %0, %1, %2;" : "=l"(su) : "l"(x0), "l"(x1));
              asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(di) : "l"(x0), "l"(x1));
              f2_unpack(su, acc[b][tl][2 * p], acc[b][tl][2 * p + 1]);
              f2_unpack(di, acc[b + s][tl][2 * p], acc[b + s][tl][2 * p + 1]);
            }
          }
        }
      }

      // epilogue: (* what is left of the scale) -> fp16 (the reference's rounding point) -> y / codes
#pragma unroll
      for (int b = 0; b < BL; ++b) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          float c0 = acc[b][tl][0], c1 = acc[b][tl][1], c2 = acc[b][tl][2], c3 = acc[b][tl][3];
          if (mul_rest) {
            f2_unpack(f2_mul(f2_pack(c0, c1), rest2), c0, c1);
            f2_unpack(f2_mul(f2_pack(c2, c3), rest2), c2, c3);
          }
          // element order inside the word: c0, c2, c1, c3
          const uint32_t ha = h2_bits(c0, c2), hb = h2_bits(c1, c3);
          const int col = b * 256 + 128 * tl;  // + group and word offsets, already in qrow / yrow
          if (want_y) *reinterpret_cast<uint2*>(yrow + col) = make_uint2(ha, hb);
          if (want_q) {
            const uint32_t codes0 = NARROW ? quant4_narrow(ha, hb, qs.n) : quant4_exact(ha, hb, qs.s);  // (division: rare scales only)
            uint32_t codes = codes0;
            if (b == 0 && tl == 0 && split && ch == 0 && lane == 0) {
              a.x0_out[m] = __ushort_as_half((unsigned short)(ha & 0xffffu));
              codes &= 0xffffff00u;
            }
            *reinterpret_cast<uint32_t*>(qrow + col) = codes;
          }
        }
      }
      if (want_q) qrow += BL * 256;
      if (want_y) yrow += BL * 256;
    }
  }
}

template <int NB, int FB, int MODE>
__global__ void __launch_bounds__(WARPS * 32, Traits<NB, FB>::MIN_CTAS) fwht_mma_kernel(const HadArgs a) {
  bool narrow = true;
  if (MODE & 1) {
    const float s0 = __ldg(a.scales), s1 = __ldg(a.scales + 1);
    narrow = ((__float_as_uint(s0) | __float_as_uint(s1)) & 0x1fffu) == 0u;  // (make_qscale's test)
  }
  if (narrow)
    fwht_mma_body<NB, FB, MODE, true>(a);
  else
    fwht_mma_body<NB, FB, MODE, false>(a);
}

template <int NB, int FB, int MODE>
static int launch_mode(const HadArgs& a, cudaStream_t st) {
  using T = Traits<NB, FB>;
  const size_t smem = (size_t)WARPS * (T::NBUF * T::N * 2) + WARPS * 16;
  MQ_DYN_SMEM((fwht_mma_kernel<NB, FB, MODE>), smem);
  int64_t blocks = (a.M + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * T::MIN_CTAS;
  if (blocks > cap) blocks = cap;
  fwht_mma_kernel<NB, FB, MODE><<<(int)blocks, WARPS * 32, smem, st>>>(a);
  return MQ_OK;
}
template <int NB, int FB>
static int launch(const HadArgs& a, cudaStream_t st) {
  if (a.q != nullptr && a.y != nullptr) return launch_mode<NB, FB, 3>(a, st);
  return a.q != nullptr ? launch_mode<NB, FB, 1>(a, st) : launch_mode<NB, FB, 2>(a, st);
}

}  // namespace hadmma

bool fwht_mma_supported(const HadArgs& a) {
  if (a.had_k != 1) return false;
  if (a.n != 1024 && a.n != 2048 && a.n != 4096 && a.n != 8192) return false;
  if (a.n_in % 8 != 0 || a.ldx % 8 != 0) return false;
  return forced_hadamard_kernel() != 1;
}

int launch_fwht_mma(const HadArgs& a, cudaStream_t st) {
  switch (a.n) {
    case 1024: return hadmma::launch<4, 0>(a, st);
    case 2048: return hadmma::launch<8, 1>(a, st);
    case 4096: return hadmma::launch<16, 0>(a, st);  // (folding block bits into stage 1 does not pay here)
    default: return hadmma::launch<32, 2>(a, st);
  }
}

}  // namespace mq
