// Stage 5: sample adaptive offset (H.265 8.7.3): deblocked picture (recon arena) -> final arena.
// Absent from the reference (slice.rs:249-251 is todo!()).  A thread produces eight
// horizontally adjacent samples from at most three rows of the deblocked picture; the rows above and below
// come out of L1/L2 because neighbouring threads just fetched them.
#include <cuda_runtime.h>

#include "kernels.h"
#include "sao_common.cuh"

namespace heic {
namespace dev {

namespace {

using sao::sao8;

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

// Sparse form for the RGB path: only the CTBs whose SAO type is non-zero (a few per cent in real streams) are processed,
// recon arena -> final arena at those CTBs; the colour kernel then reads each CTB's component from whichever arena holds
// its final samples.  One CTA per CTB row of a tile, a warp per CTB (round-robin); a warp whose CTB has SAO off leaves after
// one 16-byte load of the CTB's parameter words.
__global__ void __launch_bounds__(128) sao_sparse_kernel(Arenas A, uint32_t max_hctb) {
  const uint32_t tile = blockIdx.x / max_hctb, ry = blockIdx.x % max_hctb;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0 || ry >= (uint32_t)pp->hctb) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_comp = pp->chroma ? 3 : 1;
  for (int rx = warp; rx < pp->wctb; rx += 4) {
    const uint4 sp = *reinterpret_cast<const uint4*>(A.sao + tp->sao_off + (size_t)(ry * pp->wctb + rx) * 4);
    if (!((sp.x | sp.y | sp.z) & 3u)) continue;
#pragma unroll 1
    for (int c = 0; c < n_comp; c++) {
      const uint32_t word = c == 0 ? sp.x : (c == 1 ? sp.y : sp.z);
      if (!(word & 3u)) continue;
      const int sub = c ? 1 : 0;
      const int cs = 1 << (pp->log2_ctb - sub), pw = pp->w >> sub, ph = pp->h >> sub, pitch = c ? pp->pitch_c : pp->pitch_y;
      const int x_ctb = rx * cs, y_ctb = (int)ry * cs;
      const uint8_t* src = A.recon + tp->plane_off[c];
      uint8_t* dst = A.final_ + tp->plane_off[c];
      const int gpr = cs >> 3;  // 8-sample groups per CTB row
#pragma unroll 1
      for (int it = lane; it < gpr * cs; it += 32) {
        const int x = x_ctb + (it % gpr) * 8, y = y_ctb + it / gpr;
        if (x >= pw || y >= ph) continue;
        const uint8_t* row = src + (size_t)y * pitch;
        const uint2 in = *reinterpret_cast<const uint2*>(row + x);  // pitches are multiples of 64: inside the row
        uint32_t o[2] = {in.x, in.y};
        sao8(o, row, x, y, pw, ph, pitch, word, (int)(word & 3u));
        uint8_t* out = dst + (size_t)y * pitch + x;
        *reinterpret_cast<uint32_t*>(out) = o[0];
        if (x + 4 < pw) *reinterpret_cast<uint32_t*>(out + 4) = o[1];
      }
    }
  }
}

}  // namespace

cudaError_t launch_sao_sparse(const Arenas& A, int max_hctb, cudaStream_t stream) {
  if (!A.n_tiles || max_hctb <= 0) return cudaSuccess;
  sao_sparse_kernel<<<A.n_tiles * (uint32_t)max_hctb, 128, 0, stream>>>(A, (uint32_t)max_hctb);
  return cudaGetLastError();
}

cudaError_t launch_sao(const Arenas& A, uint32_t max_pitch, uint32_t max_h, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const uint32_t groups = ((max_pitch + 15) / 16) * max_h * 2;  // upper bound on the 16-sample groups of a tile
  const uint32_t bpt = (groups + 255) / 256;
  sao_kernel<<<A.n_tiles * bpt, 256, 0, stream>>>(A, bpt);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
