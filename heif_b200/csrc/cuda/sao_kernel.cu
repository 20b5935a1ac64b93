// Stage 5: sample adaptive offset (H.265 8.7.3): deblocked picture (recon arena) -> final arena.
// Absent from the reference (slice.rs:249-251 is todo!()).  A thread produces eight
// horizontally adjacent samples from at most three rows of the deblocked picture; the rows above and below
// come out of L1/L2 because neighbouring threads just fetched them.
#include <cuda_runtime.h>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

__device__ __forceinline__ int clip8(int v) { return min(255, max(0, v)); }
__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }

// A thread produces eight horizontally adjacent samples.  Rows are fetched as 8-byte words plus the two bytes beside
// them; sample i of the thread sits at byte i + 1 of the 10-byte window (lo = bytes 0..7, hi = bytes 8..9).
struct Window {
  uint64_t lo;
  uint32_t hi;
  __device__ __forceinline__ int at(int i) const { return i < 8 ? (int)((lo >> (8 * i)) & 0xffu) : (int)((hi >> (8 * (i - 8))) & 0xffu); }
};
__device__ __forceinline__ Window row_window(const uint8_t* row, int x, int pw) {
  const uint2 c = *reinterpret_cast<const uint2*>(row + x);
  const uint64_t c64 = ((uint64_t)c.y << 32) | c.x;
  const uint32_t l = x > 0 ? row[x - 1] : 0u;
  const uint32_t r = x + 8 < pw ? row[x + 8] : 0u;
  Window w;
  w.lo = (c64 << 8) | l;
  w.hi = (uint32_t)(c64 >> 56) | (r << 8);
  return w;
}

// SAO of eight samples at (x, y) of one plane; `out` holds the deblocked samples on entry.
__device__ __forceinline__ void sao8(uint32_t (&out)[2], const uint8_t* row, int x, int y, int pw, int ph, int pitch, uint32_t word,
                                     int type) {
  if (type) {
    // the four offsets as one word of signed nibbles; entry 0 of the edge table (edgeIdx 2 -> 0) is zero
    const uint32_t offs = (word >> 8) & 0xffffu;
    if (type == 1) {  // band offset
      const int band_pos = (int)((word >> 2) & 31u);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          int v = (int)((out[h] >> (8 * i)) & 0xffu);
          const int k = ((v >> 3) - band_pos) & 31;
          if (k < 4) v = clip8(v + ((int)(offs << (28 - 4 * k)) >> 28));
          o |= (uint32_t)v << (8 * i);
        }
        out[h] = o;
      }
    } else {  // edge offset
      const int cl = (int)((word >> 2) & 3u);
      const int dxa = cl == 1 ? 0 : (cl == 3 ? 1 : -1), dya = cl == 0 ? 0 : -1;  // b is the mirror of a
      const bool rows_ok = dya == 0 || (y > 0 && y + 1 < ph);
      if (rows_ok) {
        const Window wa = row_window(row + (ptrdiff_t)dya * pitch, x, pw);
        const Window wb = row_window(row - (ptrdiff_t)dya * pitch, x, pw);
        // edgeIdx = 2 + sign(v - a) + sign(v - b) -> offset index {1, 2, 0, 3, 4}: nibble table, 4 bits per edgeIdx
        const uint32_t table = (offs & 0xfu) | (((offs >> 4) & 0xfu) << 4) | (((offs >> 8) & 0xfu) << 12) | (((offs >> 12) & 0xfu) << 16);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          uint32_t o = 0;
#pragma unroll
          for (int i4 = 0; i4 < 4; i4++) {
            const int i = 4 * h + i4;
            int v = (int)((out[h] >> (8 * i4)) & 0xffu);
            const int xa = x + i + dxa, xb = x + i - dxa;
            if (x + i < pw && xa >= 0 && xa < pw && xb >= 0 && xb < pw) {
              const int a = wa.at(i + 1 + dxa), b = wb.at(i + 1 - dxa);
              const int e = 2 + sgn(v - a) + sgn(v - b);
              v = clip8(v + ((int)(table << (28 - 4 * e)) >> 28));
            }
            o |= (uint32_t)v << (8 * i4);
          }
          out[h] = o;
        }
      }
    }
  }
}

// grid: flat over (tile, block of 16-sample groups of the tile); the groups of the three planes are numbered consecutively,
// row by row, so CTAs stay full whatever the picture width is.  A thread moves 16 bytes; with SAO off for the CTB (the
// common case in real streams) that is a plain 16-byte copy.
__global__ void __launch_bounds__(256) sao_kernel(Arenas A, uint32_t blocks_per_tile) {
  const uint32_t tile = blockIdx.x / blocks_per_tile;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0) return;
  uint32_t g = (blockIdx.x % blocks_per_tile) * blockDim.x + threadIdx.x;
  const uint32_t gy = (uint32_t)(pp->w + 15) >> 4, gc = (uint32_t)((pp->w >> 1) + 15) >> 4;  // groups per row
  const uint32_t n_y = gy * (uint32_t)pp->h, n_c = pp->chroma ? gc * (uint32_t)(pp->h >> 1) : 0u;
  int cidx = 0;
  if (g >= n_y) {
    g -= n_y;
    cidx = 1;
    if (g >= n_c) {
      g -= n_c;
      cidx = 2;
      if (g >= n_c) return;
    }
  }
  const uint32_t gpr = cidx ? gc : gy;
  const int y = (int)(g / gpr);
  const int sub = cidx ? 1 : 0;
  const int pw = pp->w >> sub, ph = pp->h >> sub, pitch = cidx ? pp->pitch_c : pp->pitch_y;
  const int x = (int)(g % gpr) * 16;  // plane widths are multiples of 4, pitches of 64: the 16-byte access stays inside the row
  const uint8_t* src = A.recon + tp->plane_off[cidx];
  uint8_t* dst = A.final_ + tp->plane_off[cidx];
  const uint8_t* row = src + (size_t)y * pitch;
  const int log2_cs = pp->log2_ctb - sub;
  const bool enabled = cidx == 0 ? tp->sao_luma : tp->sao_chroma;
  const uint32_t* sao_row = A.sao + tp->sao_off + (size_t)((y >> log2_cs) * pp->wctb) * 4 + cidx;
  const uint4 center = *reinterpret_cast<const uint4*>(row + x);
  uint32_t lo[2] = {center.x, center.y}, hi[2] = {center.z, center.w};
  if (enabled) {
    const uint32_t w0 = sao_row[(size_t)(x >> log2_cs) * 4];
    if (w0 & 3u) sao8(lo, row, x, y, pw, ph, pitch, w0, (int)(w0 & 3u));
    if (x + 8 < pw) {
      const uint32_t w1 = sao_row[(size_t)((x + 8) >> log2_cs) * 4];
      if (w1 & 3u) sao8(hi, row, x + 8, y, pw, ph, pitch, w1, (int)(w1 & 3u));
    }
  }
  uint8_t* o = dst + (size_t)y * pitch + x;
  if (x + 16 <= pw) {
    *reinterpret_cast<uint4*>(o) = make_uint4(lo[0], lo[1], hi[0], hi[1]);
  } else {  // ragged right edge: 4, 8 or 12 valid samples
    const int n = pw - x;
    *reinterpret_cast<uint32_t*>(o) = lo[0];
    if (n > 4) *reinterpret_cast<uint32_t*>(o + 4) = lo[1];
    if (n > 8) *reinterpret_cast<uint32_t*>(o + 8) = hi[0];
  }
}

}  // namespace

cudaError_t launch_sao(const Arenas& A, uint32_t max_pitch, uint32_t max_h, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const uint32_t groups = ((max_pitch + 15) / 16) * max_h * 2;  // upper bound on the 16-sample groups of a tile
  const uint32_t bpt = (groups + 255) / 256;
  sao_kernel<<<A.n_tiles * bpt, 256, 0, stream>>>(A, bpt);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
