// W4A8 GEMM for decode-sized batches (M <= 16): weight streaming on CUDA cores.
//
// At M <= 16 the layer is bound by the N*K/2-byte weight stream (K=4096, N=11008
// is 23 MB -> 3.4 us at the measured 6.55 TB/s), not by the tensor pipe, and the 128 x 256 tiles of
// the tcgen05 kernels leave only N/256 = 16..43 CTAs to pull it.  This kernel splits K across
// CTAs so that >= 2 CTAs per SM stream concurrently:
//   grid = (n-blocks, k-splits); CTA = 256 threads = the 256 rows of one packed weight tile column
//   thread r: per k-block 4 coalesced 128-bit loads (the tile-major layout puts row r's 16-byte
//             chunks of consecutive threads side by side), nibble split with two logic ops per 8
//             weights (operands are 16 x code, exactly as in the tensor-core kernels), 32 x M
//             dp4a against the activation rows broadcast from shared memory
//   partial sums are merged with int32 atomics in a caller-provided, self-cleaning workspace
//   (integer adds are exact and order independent); the last CTA of an n-block applies the same
//   epilogue arithmetic as mq_w4a8_gemm -> bit-identical results.
#include "mq_gemm_common.cuh"

namespace mq {
namespace gemv {

constexpr int MAXM = 16;
constexpr int THREADS = 256;
constexpr int DEPTH = 3;

struct Params {
  const int8_t* qa;
  int64_t lda;
  gemm::Params g;
  int* ws_acc;       // [MAXM][n_tiles*256] int32, zero on entry, zero on exit
  unsigned* ws_cnt;  // [n_tiles] arrival counters, zero on entry, zero on exit
  int ksplit, kb_per_split;
};

template <int M>
__global__ void __launch_bounds__(THREADS)
w4a8_gemv_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];  // A slab: [M][kb_per_split*128] int8
  const gemm::Params& g = p.g;
  const int nb = blockIdx.x, ks = blockIdx.y;
  const int r = threadIdx.x;
  const int kb0 = ks * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, g.k_tiles);
  const int slab = (kb1 - kb0) * 128;  // bytes per activation row in this k range
  // stage the activation slab (16-byte vectors, coalesced)
  for (int i = threadIdx.x; i < M * (slab >> 4); i += THREADS) {
    const int m = i / (slab >> 4), c = i - m * (slab >> 4);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (m < g.M) v = *reinterpret_cast<const uint4*>(p.qa + (int64_t)m * p.lda + kb0 * 128 + c * 16);
    *reinterpret_cast<uint4*>(smem + m * slab + c * 16) = v;
  }
  __syncthreads();
  int acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0;
  const uint8_t* wrow = g.w4 + w4_blob_offset(nb, kb0, g.k_tiles) + (size_t)r * 16;
  // DEPTH k-blocks of weights live in registers: while one is consumed, DEPTH-1 are in flight
  uint4 buf[DEPTH][4];
#pragma unroll
  for (int d = 0; d < DEPTH; ++d)
    if (kb0 + d < kb1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) buf[d][j] = ldg_stream(wrow + (size_t)d * kW4BlobBytes + (size_t)j * kW4TileN * 16);
    }
  for (int kbase = kb0; kbase < kb1; kbase += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const int kb = kbase + d;
      if (kb < kb1) {
        const uint8_t* arow = smem + (kb - kb0) * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w[4] = {buf[d][j].x, buf[d][j].y, buf[d][j].z, buf[d][j].w};
          int lo[4], hi[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            lo[i] = (int)((w[i] << 4) & 0xF0F0F0F0u);  // 16 x code, k = 32j + 4i .. +3
            hi[i] = (int)(w[i] & 0xF0F0F0F0u);         // 16 x code, k = 32j + 16 + 4i .. +3
          }
#pragma unroll
          for (int m = 0; m < M; ++m) {
            // 16 + 16 activations of row m: two broadcast 128-bit shared loads feed 8 dp4a
            const int4 a_lo = *reinterpret_cast<const int4*>(arow + m * slab + 32 * j);
            const int4 a_hi = *reinterpret_cast<const int4*>(arow + m * slab + 32 * j + 16);
            int t = acc[m];
            t = __dp4a(a_lo.x, lo[0], t);
            t = __dp4a(a_lo.y, lo[1], t);
            t = __dp4a(a_lo.z, lo[2], t);
            t = __dp4a(a_lo.w, lo[3], t);
            t = __dp4a(a_hi.x, hi[0], t);
            t = __dp4a(a_hi.y, hi[1], t);
            t = __dp4a(a_hi.z, hi[2], t);
            t = __dp4a(a_hi.w, hi[3], t);
            acc[m] = t;
          }
        }
        if (kb + DEPTH < kb1) {
          const uint8_t* wn = wrow + (size_t)(kb + DEPTH - kb0) * kW4BlobBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) buf[d][j] = ldg_stream(wn + (size_t)j * kW4TileN * 16);
        }
      }
    }
  }
  // merge the k-split partial sums
  const int ncols = g.n_tiles * kW4TileN;
  int* wsrow = p.ws_acc + nb * kW4TileN + r;
  if (p.ksplit > 1) {
#pragma unroll
    for (int m = 0; m < M; ++m)
      if (acc[m] != 0) atomicAdd(wsrow + m * ncols, acc[m]);
    __threadfence();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) last = (atomicAdd(p.ws_cnt + nb, 1u) == (unsigned)(p.ksplit - 1));
    __syncthreads();
    if (!last) return;
    __threadfence();
#pragma unroll
    for (int m = 0; m < M; ++m) {
      acc[m] = __ldcg(wsrow + m * ncols);
      wsrow[m * ncols] = 0;  // leave the workspace clean for the next call
    }
    if (threadIdx.x == 0) p.ws_cnt[nb] = 0;
  }
  // epilogue: identical arithmetic to the tensor-core kernels (operands carried 16 x code)
  const int n = nb * kW4TileN + r;
  if (n >= g.N) return;
  const float sw16 = __fmul_rn(g.s_w[n], 0.0625f);
  const float bias = g.bias != nullptr ? __half2float(g.bias[n]) : 0.f;
  const float w0 = g.w0 != nullptr ? __half2float(g.w0[n]) : 0.f;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    if (m < g.M) {
      const float sa = g.s_a_per_row ? g.s_a[m] : g.s_a[0];
      float y = __fmul_rn((float)acc[m], __fmul_rn(sa, sw16));
      if (g.x0 != nullptr) y = __fadd_rn(y, __fmul_rn(__half2float(g.x0[m]), w0));
      if (g.bias != nullptr) y = __fadd_rn(y, bias);
      g.out[(int64_t)m * g.ldo + n] = __float2half_rn(y);
    }
  }
}

template <int M>
static int launch(const Params& p, dim3 grid, size_t smem, cudaStream_t st) {
  MQ_DYN_SMEM(w4a8_gemv_kernel<M>, smem);
  w4a8_gemv_kernel<M><<<grid, THREADS, smem, st>>>(p);
  return MQ_OK;
}

}  // namespace gemv

size_t gemv_workspace_bytes(int64_t N) {
  const int64_t n_tiles = (N + kW4TileN - 1) / kW4TileN;
  return (size_t)(gemv::MAXM * n_tiles * kW4TileN * sizeof(int) + n_tiles * sizeof(unsigned));
}

int launch_w4a8_gemv(const int8_t* qa, int64_t lda, const gemm::Params& g, void* workspace, cudaStream_t st) {
  using namespace gemv;
  Params p;
  p.qa = qa;
  p.lda = lda;
  p.g = g;
  const int n_tiles = g.n_tiles;
  p.ws_acc = reinterpret_cast<int*>(workspace);
  p.ws_cnt = reinterpret_cast<unsigned*>(p.ws_acc + (size_t)MAXM * n_tiles * kW4TileN);
  // aim at >= 2 CTAs per SM; keep at least 4 k-blocks per CTA so that the prefetch has depth
  int ksplit = (2 * num_sms() + n_tiles - 1) / n_tiles;
  const int max_split = g.k_tiles / 4 > 0 ? g.k_tiles / 4 : 1;
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  p.kb_per_split = (g.k_tiles + ksplit - 1) / ksplit;
  p.ksplit = (g.k_tiles + p.kb_per_split - 1) / p.kb_per_split;
  const int Mr = g.M <= 1 ? 1 : g.M <= 2 ? 2 : g.M <= 4 ? 4 : g.M <= 8 ? 8 : 16;
  const size_t smem = (size_t)Mr * p.kb_per_split * 128;
  const dim3 grid(n_tiles, p.ksplit);
  switch (Mr) {
    case 1: return launch<1>(p, grid, smem, st);
    case 2: return launch<2>(p, grid, smem, st);
    case 4: return launch<4>(p, grid, smem, st);
    case 8: return launch<8>(p, grid, smem, st);
    default: return launch<16>(p, grid, smem, st);
  }
}

}  // namespace mq
