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

// A thread converts a 4 x 2 block of the canvas (two rows share one row of chroma).
// grid: flat over (image, pair of canvas rows, block of 4-pixel groups).
__global__ void __launch_bounds__(256) color_stitch_kernel(ColorJob J, Coeffs K, uint32_t row_pairs, uint32_t xblocks) {
  const uint32_t per_image = row_pairs * xblocks;
  const uint32_t image = blockIdx.x / per_image, rem = blockIdx.x % per_image;
  const uint32_t y = (rem / xblocks) * 2;
  const uint32_t x = ((rem % xblocks) * blockDim.x + threadIdx.x) * 4;
  if (x >= J.out_w || y >= J.out_h) return;
  // all four pixels lie in one tile: tile widths are multiples of 8
  const uint32_t tc = x / J.tile_w, tr = y / J.tile_h, lx = x - tc * J.tile_w, ly = y - tr * J.tile_h;
  const uint8_t* t = J.planes + (size_t)(image * J.grid_cols * J.grid_rows + tr * J.grid_cols + tc) * J.tile_stride;
  const uint32_t y0 = *reinterpret_cast<const uint32_t*>(t + (size_t)ly * J.pitch_y + lx);
  const bool two_rows = y + 1 < J.out_h;  // tile heights are even, so row y + 1 is in the same tile
  const uint32_t y1 = two_rows ? *reinterpret_cast<const uint32_t*>(t + (size_t)(ly + 1) * J.pitch_y + lx) : 0u;
  uint32_t cb = 0x8080u, cr = 0x8080u;
  if (J.chroma) {
    const size_t co = (size_t)(ly >> 1) * J.pitch_c + (lx >> 1);
    cb = *reinterpret_cast<const uint16_t*>(t + J.cb_off + co);
    cr = *reinterpret_cast<const uint16_t*>(t + J.cr_off + co);
  }
  uint32_t px[2][4][3];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = (int)((cb >> (8 * (i >> 1))) & 0xffu) - 128, d = (int)((cr >> (8 * (i >> 1))) & 0xffu) - 128;
    const int r_add = K.rv * d + 128, g_add = K.gu * c + K.gv * d + 128, b_add = K.bu * c + 128;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int yy = K.y_mul * ((int)(((r ? y1 : y0) >> (8 * i)) & 0xffu) - K.y_sub);
      px[r][i][0] = clip8((yy + r_add) >> 8);
      px[r][i][1] = clip8((yy + g_add) >> 8);
      px[r][i][2] = clip8((yy + b_add) >> 8);
    }
  }
  uint8_t* img = J.rgb + (size_t)image * J.image_stride;
  if (J.rotation == 0) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r == 1 && !two_rows) break;
      uint8_t* o = img + (size_t)(y + r) * J.pitch + (size_t)x * 3;
      if (x + 4 <= J.out_w && (((uintptr_t)o) & 3u) == 0) {
        uint32_t w0 = px[r][0][0] | (px[r][0][1] << 8) | (px[r][0][2] << 16) | (px[r][1][0] << 24);
        uint32_t w1 = px[r][1][1] | (px[r][1][2] << 8) | (px[r][2][0] << 16) | (px[r][2][1] << 24);
        uint32_t w2 = px[r][2][2] | (px[r][3][0] << 8) | (px[r][3][1] << 16) | (px[r][3][2] << 24);
        reinterpret_cast<uint32_t*>(o)[0] = w0;
        reinterpret_cast<uint32_t*>(o)[1] = w1;
        reinterpret_cast<uint32_t*>(o)[2] = w2;
      } else {
        for (int i = 0; i < 4 && x + i < J.out_w; i++)
          for (int ch = 0; ch < 3; ch++) o[i * 3 + ch] = (uint8_t)px[r][i][ch];
      }
    }
  } else {
    // irot: anti-clockwise quarter turns (ISO/IEC 23008-12 6.5.10); canvas (x, y) -> rotated position
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r == 1 && !two_rows) break;
      for (int i = 0; i < 4 && x + i < J.out_w; i++) {
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
  const uint32_t groups = (job.out_w + 3) / 4;
  const uint32_t xblocks = (groups + 255) / 256, row_pairs = (job.out_h + 1) / 2;
  color_stitch_kernel<<<job.n_images * row_pairs * xblocks, 256, 0, stream>>>(job, k, row_pairs, xblocks);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
