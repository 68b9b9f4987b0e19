// Packed INT4 weight layout shared by the pack kernels and the W4A8 GEMM.
//
// The [N, K] code matrix is cut into tiles of 256 output channels x 128 input channels.
// Tile (nb, kb) is one contiguous 16 KB blob at byte offset (nb * (K/128) + kb) * 16384, so a
// CTA streaming one n-block along K reads one contiguous region.  Inside a blob:
//
//   16-byte chunk  q = j * 256 + r      j = 0..3 (32-wide k group), r = 0..255 (row in tile)
//   byte b of the chunk (b = 0..15):    low  nibble = code[nb*256 + r][kb*128 + 32*j + b]
//                                       high nibble = code[nb*256 + r][kb*128 + 32*j + 16 + b]
//
// Chunk-interleaving (r fastest) makes a warp's 128-bit loads land on 512 contiguous bytes;
// the nibble split means the in-kernel unpack is two logic ops per 8 weights and yields two
// 16-byte runs of CONSECUTIVE k:
//   (word << 4) & 0xF0F0F0F0  -> 16 * code[k .. k+3]      (low nibbles)
//    word       & 0xF0F0F0F0  -> 16 * code[k+16 .. k+19]  (high nibbles)
// i.e. the int8 operand holds 16x the code (exact, |16*code| <= 128) and the epilogue folds the
// 1/16 into the scale -- no sign-extension arithmetic in the main loop.
// Rows N..round_up(N,256) are zero.
#pragma once
#include <stdint.h>

namespace mq {
constexpr int kW4TileN = 256;
constexpr int kW4TileK = 128;
constexpr int kW4BlobBytes = kW4TileN * kW4TileK / 2;  // 16384

__host__ __device__ inline int64_t w4_blob_offset(int64_t nb, int64_t kb, int64_t k_tiles) {
  return (nb * k_tiles + kb) * (int64_t)kW4BlobBytes;
}
__host__ __device__ inline int w4_chunk_offset(int j, int r) { return (j * kW4TileN + r) * 16; }
}  // namespace mq
