// Stage 6: YCbCr 4:2:0 -> RGB8 fused with HEIF grid stitching, conformance/canvas crop and (optionally)
// the irot rotation — one bandwidth-bound pass: 1.5 B read + 3 B written per output pixel.
//
// The reference's src/color is an ICC-header stub, not a converter, so the conversion is the
// frozen integer definition of SURVEY row C1 (libheif-style nearest-neighbour chroma, 8.8 fixed point):
//   full range BT.601:  R = clip8(Y + ((359 d + 128) >> 8)),  G = clip8(Y + ((-88 c - 183 d + 128) >> 8)),
//                       B = clip8(Y + ((454 c + 128) >> 8)),  c = Cb - 128, d = Cr - 128
// with the BT.709 / limited-range variants listed in oracle/hevc_oracle.c (color_coeffs).
#include <cuda_runtime.h>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

struct Coeffs {
  int y_mul, y_sub, rv, gu, gv, bu;
};

__device__ __forceinline__ uint32_t clip8(int v) { return (uint32_t)min(255, max(0, v)); }

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// A thread converts an 8 x 2 block of the canvas (two rows share one row of chroma): 16 + 8 bytes in, 48 bytes out.
// FULL: full-range input, where (256 Y + t) >> 8 == Y + (t >> 8), so the chroma terms are computed once per 2x2 quad
// and a pixel costs an add and a clamp per channel.
// grid: flat over (image, pair of canvas rows, block of 8-pixel groups).
template <bool FULL>
__global__ void __launch_bounds__(256) color_stitch_kernel(ColorJob J, Coeffs K, uint32_t row_pairs, uint32_t xblocks) {
  const uint32_t per_image = row_pairs * xblocks;
  const uint32_t image = blockIdx.x / per_image, rem = blockIdx.x % per_image;
  const uint32_t y = (rem / xblocks) * 2;
  const uint32_t x = ((rem % xblocks) * blockDim.x + threadIdx.x) * 8;
  if (x >= J.out_w || y >= J.out_h) return;
  // all eight pixels lie in one tile: tile widths are multiples of 8
  const uint32_t tc = x / J.tile_w, tr = y / J.tile_h, lx = x - tc * J.tile_w, ly = y - tr * J.tile_h;
  const uint8_t* t = J.planes + (size_t)(image * J.grid_cols * J.grid_rows + tr * J.grid_cols + tc) * J.tile_stride;
  const bool two_rows = y + 1 < J.out_h;  // tile heights are even, so row y + 1 is in the same tile
  uint2 yr[2];
  yr[0] = *reinterpret_cast<const uint2*>(t + (size_t)ly * J.pitch_y + lx);
  yr[1] = two_rows ? *reinterpret_cast<const uint2*>(t + (size_t)(ly + 1) * J.pitch_y + lx) : make_uint2(0u, 0u);
  uint32_t cb = 0x80808080u, cr = 0x80808080u;
  if (J.chroma) {
    const size_t co = (size_t)(ly >> 1) * J.pitch_c + (lx >> 1);
    cb = *reinterpret_cast<const uint32_t*>(t + J.cb_off + co);
    cr = *reinterpret_cast<const uint32_t*>(t + J.cr_off + co);
  }
  uint32_t px[2][8][3];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = (int)((cb >> (8 * j)) & 0xffu) - 128, d = (int)((cr >> (8 * j)) & 0xffu) - 128;
    const int r_add = K.rv * d + 128, g_add = K.gu * c + K.gv * d + 128, b_add = K.bu * c + 128;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int p = 2 * j + i;
        const int Y = (int)(((p < 4 ? yr[r].x : yr[r].y) >> (8 * (p & 3))) & 0xffu);
        if (FULL) {
          px[r][p][0] = clip8(Y + (r_add >> 8));
          px[r][p][1] = clip8(Y + (g_add >> 8));
          px[r][p][2] = clip8(Y + (b_add >> 8));
        } else {
          const int yy = K.y_mul * (Y - K.y_sub);
          px[r][p][0] = clip8((yy + r_add) >> 8);
          px[r][p][1] = clip8((yy + g_add) >> 8);
          px[r][p][2] = clip8((yy + b_add) >> 8);
        }
      }
  }
  uint8_t* img = J.rgb + (size_t)image * J.image_stride;
  if (J.rotation == 0) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r == 1 && !two_rows) break;
      uint8_t* o = img + (size_t)(y + r) * J.pitch + (size_t)x * 3;
      if (x + 8 <= J.out_w && (((uintptr_t)o) & 7u) == 0) {
        uint32_t w[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {  // bytes 4k .. 4k+3 of R0 G0 B0 R1 G1 B1 ...
          const int b0 = 4 * k, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
          w[k] = pack4(px[r][b0 / 3][b0 % 3], px[r][b1 / 3][b1 % 3], px[r][b2 / 3][b2 % 3], px[r][b3 / 3][b3 % 3]);
        }
        reinterpret_cast<uint2*>(o)[0] = make_uint2(w[0], w[1]);
        reinterpret_cast<uint2*>(o)[1] = make_uint2(w[2], w[3]);
        reinterpret_cast<uint2*>(o)[2] = make_uint2(w[4], w[5]);
      } else {
        for (int i = 0; i < 8 && x + i < J.out_w; i++)
          for (int ch = 0; ch < 3; ch++) o[i * 3 + ch] = (uint8_t)px[r][i][ch];
      }
    }
  } else {
    // irot: anti-clockwise quarter turns (ISO/IEC 23008-12 6.5.10); canvas (x, y) -> rotated position
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r == 1 && !two_rows) break;
      for (int i = 0; i < 8 && x + i < J.out_w; i++) {
        const uint32_t sx = x + i, sy = y + r;
        uint32_t dx, dy;
        if (J.rotation == 1) {
          dx = sy;
          dy = J.out_w - 1 - sx;
        } else if (J.rotation == 2) {
          dx = J.out_w - 1 - sx;
          dy = J.out_h - 1 - sy;
        } else {
          dx = J.out_h - 1 - sy;
          dy = sx;
        }
        uint8_t* o = img + (size_t)dy * J.pitch + (size_t)dx * 3;
        o[0] = (uint8_t)px[r][i][0];
        o[1] = (uint8_t)px[r][i][1];
        o[2] = (uint8_t)px[r][i][2];
      }
    }
  }
}

}  // namespace

cudaError_t launch_color(const ColorJob& job, cudaStream_t stream) {
  if (!job.n_images || !job.out_w || !job.out_h) return cudaSuccess;
  Coeffs k;
  const bool bt709 = job.matrix_coeffs == 1;
  if (job.full_range) {
    k.y_mul = 256;
    k.y_sub = 0;
    if (bt709) k.rv = 403, k.gu = -48, k.gv = -120, k.bu = 475;
    else k.rv = 359, k.gu = -88, k.gv = -183, k.bu = 454;
  } else {
    k.y_mul = 298;
    k.y_sub = 16;
    if (bt709) k.rv = 459, k.gu = -55, k.gv = -136, k.bu = 541;
    else k.rv = 409, k.gu = -100, k.gv = -208, k.bu = 516;
  }
  const uint32_t groups = (job.out_w + 7) / 8;
  const uint32_t xblocks = (groups + 255) / 256, row_pairs = (job.out_h + 1) / 2;
  if (job.full_range) color_stitch_kernel<true><<<job.n_images * row_pairs * xblocks, 256, 0, stream>>>(job, k, row_pairs, xblocks);
  else color_stitch_kernel<false><<<job.n_images * row_pairs * xblocks, 256, 0, stream>>>(job, k, row_pairs, xblocks);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
