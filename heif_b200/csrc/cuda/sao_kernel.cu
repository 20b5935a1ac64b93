// Stage 5: sample adaptive offset (H.265 8.7.3): deblocked picture (recon arena) -> final arena.
// Absent from the reference (slice.rs:249-251 is todo!()).  Purely bandwidth bound: a thread produces four
// horizontally adjacent samples from at most three rows of the deblocked picture; the rows above and below
// come out of L1/L2 because neighbouring threads just fetched them.
#include <cuda_runtime.h>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

__device__ __forceinline__ int clip8(int v) { return min(255, max(0, v)); }
__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }

// bytes x-1 .. x+4 of a row as a 48-bit window (missing neighbours read as 0; callers mask them out)
__device__ __forceinline__ uint64_t row_window(const uint8_t* row, int x, int pw) {
  const uint32_t c = *reinterpret_cast<const uint32_t*>(row + x);
  const uint32_t l = x > 0 ? row[x - 1] : 0u;
  const uint32_t r = x + 4 < pw ? row[x + 4] : 0u;
  return (uint64_t)l | ((uint64_t)c << 8) | ((uint64_t)r << 40);
}

// grid: flat over (tile, row of the three stacked planes, block of 4-sample groups)
__global__ void __launch_bounds__(128) sao_kernel(Arenas A, uint32_t rows, uint32_t xblocks) {
  const uint32_t per_tile = rows * xblocks;
  const uint32_t tile = blockIdx.x / per_tile, rem = blockIdx.x % per_tile;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0) return;
  int y = (int)(rem / xblocks), cidx = 0;
  if (y >= pp->h) {
    if (!pp->chroma) return;
    y -= pp->h;
    cidx = 1;
    if (y >= (pp->h >> 1)) {
      y -= pp->h >> 1;
      cidx = 2;
      if (y >= (pp->h >> 1)) return;
    }
  }
  const int sub = cidx ? 1 : 0;
  const int pw = pp->w >> sub, ph = pp->h >> sub, pitch = cidx ? pp->pitch_c : pp->pitch_y;
  const int x = (int)((rem % xblocks) * blockDim.x + threadIdx.x) * 4;
  if (x >= pw) return;
  const uint8_t* src = A.recon + tp->plane_off[cidx];
  uint8_t* dst = A.final_ + tp->plane_off[cidx];
  const uint8_t* row = src + (size_t)y * pitch;
  const int log2_cs = pp->log2_ctb - sub;
  const uint32_t word = A.sao[tp->sao_off + (size_t)((y >> log2_cs) * pp->wctb + (x >> log2_cs)) * 4 + cidx];
  const bool enabled = cidx == 0 ? tp->sao_luma : tp->sao_chroma;
  const int type = enabled ? (int)(word & 3u) : 0;
  const uint32_t center = *reinterpret_cast<const uint32_t*>(row + x);
  uint32_t out = center;
  if (type) {
    int off[5];
    off[0] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) off[k + 1] = (int)(((word >> (8 + 4 * k)) & 15u) ^ 8u) - 8;
    if (type == 1) {  // band offset
      const int band_pos = (int)((word >> 2) & 31u);
      out = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int v = (int)((center >> (8 * i)) & 0xffu);
        const int k = ((v >> 3) - band_pos) & 31;
        if (k < 4) v = clip8(v + (k == 0 ? off[1] : k == 1 ? off[2] : k == 2 ? off[3] : off[4]));
        out |= (uint32_t)v << (8 * i);
      }
    } else {  // edge offset
      const int cl = (int)((word >> 2) & 3u);
      const int dxa = cl == 1 ? 0 : (cl == 3 ? 1 : -1), dya = cl == 0 ? 0 : -1;  // b is the mirror of a
      const bool rows_ok = dya == 0 || (y > 0 && y + 1 < ph);
      if (rows_ok) {
        const uint64_t wa = row_window(row + (ptrdiff_t)dya * pitch, x, pw);
        const uint64_t wb = row_window(row - (ptrdiff_t)dya * pitch, x, pw);
        out = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          int v = (int)((center >> (8 * i)) & 0xffu);
          const int xa = x + i + dxa, xb = x + i - dxa;
          if (xa >= 0 && xa < pw && xb >= 0 && xb < pw) {
            const int a = (int)((wa >> (8 * (i + 1 + dxa))) & 0xffu);
            const int b = (int)((wb >> (8 * (i + 1 - dxa))) & 0xffu);
            const int e = 2 + sgn(v - a) + sgn(v - b);
            const int o = e == 0 ? off[1] : e == 1 ? off[2] : e == 2 ? 0 : e == 3 ? off[3] : off[4];
            v = clip8(v + o);
          }
          out |= (uint32_t)v << (8 * i);
        }
      }
    }
  }
  *reinterpret_cast<uint32_t*>(dst + (size_t)y * pitch + x) = out;
}

}  // namespace

cudaError_t launch_sao(const Arenas& A, uint32_t max_pitch, uint32_t max_h, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const uint32_t xblocks = (max_pitch / 4 + 127) / 128, rows = max_h * 2;
  sao_kernel<<<A.n_tiles * rows * xblocks, 128, 0, stream>>>(A, rows, xblocks);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
