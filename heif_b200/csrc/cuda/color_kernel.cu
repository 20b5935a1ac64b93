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
#include "sao_common.cuh"

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

// SAO of one 8 x 2 luma / 4 x 1 chroma half of a thread's block.  Out of line: real streams switch SAO on for a small
// share of the CTBs, and keeping this code off the main path keeps the conversion loop compact.
struct HalfBlock {
  uint2 a, b;        // luma rows ly, ly + 1
  uint32_t cb, cr;
};
struct SaoGeom {
  uint64_t cb_off, cr_off;
  int pitch_y, pitch_c, tile_w, tile_h, chroma;
};
__device__ __noinline__ HalfBlock sao_half(SaoGeom J, const uint8_t* t, uint32_t lx, uint32_t ly, bool two_rows, uint4 sp, HalfBlock hb) {
  if (sp.x & 3u) {
    uint32_t o[2] = {hb.a.x, hb.a.y};
    sao::sao8(o, t + (size_t)ly * J.pitch_y, (int)lx, (int)ly, (int)J.tile_w, (int)J.tile_h, (int)J.pitch_y, sp.x, (int)(sp.x & 3u));
    hb.a = make_uint2(o[0], o[1]);
    if (two_rows) {
      uint32_t p[2] = {hb.b.x, hb.b.y};
      sao::sao8(p, t + (size_t)(ly + 1) * J.pitch_y, (int)lx, (int)ly + 1, (int)J.tile_w, (int)J.tile_h, (int)J.pitch_y, sp.x, (int)(sp.x & 3u));
      hb.b = make_uint2(p[0], p[1]);
    }
  }
  if (J.chroma) {
    const int cx = (int)(lx >> 1), cy = (int)(ly >> 1), x8 = cx & ~7;  // the aligned group of eight holding our four
#pragma unroll 1
    for (int comp = 1; comp <= 2; comp++) {
      const uint32_t w = comp == 1 ? sp.y : sp.z;
      if (w & 3u) {
        const uint8_t* crow = t + (comp == 1 ? J.cb_off : J.cr_off) + (size_t)cy * J.pitch_c;
        const uint2 c8 = *reinterpret_cast<const uint2*>(crow + x8);
        uint32_t o[2] = {c8.x, c8.y};
        sao::sao8(o, crow, x8, cy, (int)(J.tile_w >> 1), (int)(J.tile_h >> 1), (int)J.pitch_c, w, (int)(w & 3u));
        (comp == 1 ? hb.cb : hb.cr) = o[(cx >> 2) & 1];
      }
    }
  }
  return hb;
}

// A thread converts a 16 x 2 block of the canvas (two rows share one row of chroma): 32 + 16 bytes in, 96 bytes out, all as
// 16-byte accesses.  FULL: full-range input, where (256 Y + t) >> 8 == Y + (t >> 8), so the chroma terms are computed
// once per 2x2 quad and a pixel costs an add and a clamp per channel.
// grid: flat over (image, pair of canvas rows, block of 16-pixel groups).
// FUSED: the planes are the deblocked reconstruction and SAO is applied here (saves writing and re-reading the final planes).
template <bool FULL, bool FUSED>
__global__ void __launch_bounds__(128) color_stitch_kernel(ColorJob J, Coeffs K, uint32_t row_pairs, uint32_t xblocks) {
  const uint32_t per_image = row_pairs * xblocks;
  const uint32_t image = blockIdx.x / per_image, rem = blockIdx.x % per_image;
  const uint32_t y = (rem / xblocks) * 2;
  const uint32_t x = ((rem % xblocks) * blockDim.x + threadIdx.x) * 16;
  if (x >= J.out_w || y >= J.out_h) return;
  // tile widths are multiples of 8: the first and the second 8 pixels may lie in different tiles
  const bool two_rows = y + 1 < J.out_h;  // tile heights are even, so row y + 1 is in the same tile
  const uint32_t tr = y / J.tile_h, ly = y - tr * J.tile_h;
  uint32_t yv[2][4], cbv[2], crv[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t xh = x + 8 * h;
    if (xh < J.out_w) {
      const uint32_t tc = xh / J.tile_w, lx = xh - tc * J.tile_w;
      const uint32_t ti = image * J.grid_cols * J.grid_rows + tr * J.grid_cols + tc;
      const uint8_t* t = J.planes + (size_t)ti * J.tile_stride;
      uint2 a = *reinterpret_cast<const uint2*>(t + (size_t)ly * J.pitch_y + lx);
      uint2 b = two_rows ? *reinterpret_cast<const uint2*>(t + (size_t)(ly + 1) * J.pitch_y + lx) : make_uint2(0u, 0u);
      cbv[h] = crv[h] = 0x80808080u;
      if (J.chroma) {
        const size_t co = (size_t)(ly >> 1) * J.pitch_c + (lx >> 1);
        cbv[h] = *reinterpret_cast<const uint32_t*>(t + J.cb_off + co);
        crv[h] = *reinterpret_cast<const uint32_t*>(t + J.cr_off + co);
      }
      if (FUSED) {
        // the eight luma samples (both rows: ly is even) and the four chroma samples lie in one CTB; its parameter words
        // are zero (type 0) for a component whose slice-level SAO flag is off, so one 16-byte load decides everything
        const uint4 sp = *reinterpret_cast<const uint4*>(J.sao + (size_t)ti * J.sao_stride +
                                                         (size_t)((ly >> J.log2_ctb) * J.wctb + (lx >> J.log2_ctb)) * 4);
        if ((sp.x | sp.y | sp.z) & 3u) {
          HalfBlock hb = {a, b, cbv[h], crv[h]};
          const SaoGeom g = {J.cb_off, J.cr_off, (int)J.pitch_y, (int)J.pitch_c, (int)J.tile_w, (int)J.tile_h, (int)J.chroma};
          hb = sao_half(g, t, lx, ly, two_rows, sp, hb);
          a = hb.a, b = hb.b, cbv[h] = hb.cb, crv[h] = hb.cr;
        }
      }
      yv[0][2 * h] = a.x, yv[0][2 * h + 1] = a.y, yv[1][2 * h] = b.x, yv[1][2 * h + 1] = b.y;
    } else {
      yv[0][2 * h] = yv[0][2 * h + 1] = yv[1][2 * h] = yv[1][2 * h + 1] = 0u;
      cbv[h] = crv[h] = 0x80808080u;
    }
  }
  uint8_t* img = J.rgb + (size_t)image * J.image_stride;
  uint32_t w[2][12];  // packed RGB of the two rows
#pragma unroll
  for (int q = 0; q < 4; q++) {  // groups of four pixels -> three output words per row
    uint32_t px[2][4][3];
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const int j = 2 * q + jj;  // chroma sample index 0..7
      const int c = (int)((cbv[j >> 2] >> (8 * (j & 3))) & 0xffu) - 128, d = (int)((crv[j >> 2] >> (8 * (j & 3))) & 0xffu) - 128;
      const int r_add = K.rv * d + 128, g_add = K.gu * c + K.gv * d + 128, b_add = K.bu * c + 128;
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int p = 2 * jj + i;  // pixel within the group
          const int Y = (int)((yv[r][q] >> (8 * p)) & 0xffu);
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
#pragma unroll
    for (int r = 0; r < 2; r++) {
      w[r][3 * q + 0] = pack4(px[r][0][0], px[r][0][1], px[r][0][2], px[r][1][0]);
      w[r][3 * q + 1] = pack4(px[r][1][1], px[r][1][2], px[r][2][0], px[r][2][1]);
      w[r][3 * q + 2] = pack4(px[r][2][2], px[r][3][0], px[r][3][1], px[r][3][2]);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    if (r == 1 && !two_rows) break;
    if (J.rotation == 0) {
      uint8_t* o = img + (size_t)(y + r) * J.pitch + (size_t)x * 3;
      if (x + 16 <= J.out_w && (((uintptr_t)o) & 15u) == 0) {
        reinterpret_cast<uint4*>(o)[0] = make_uint4(w[r][0], w[r][1], w[r][2], w[r][3]);
        reinterpret_cast<uint4*>(o)[1] = make_uint4(w[r][4], w[r][5], w[r][6], w[r][7]);
        reinterpret_cast<uint4*>(o)[2] = make_uint4(w[r][8], w[r][9], w[r][10], w[r][11]);
      } else {
        for (int i = 0; i < 16 && x + i < J.out_w; i++)
          for (int ch = 0; ch < 3; ch++) {
            const int bidx = i * 3 + ch;
            o[bidx] = (uint8_t)(w[r][bidx >> 2] >> (8 * (bidx & 3)));
          }
      }
    } else {
      // irot: anti-clockwise quarter turns (ISO/IEC 23008-12 6.5.10); canvas (x, y) -> rotated position
      for (int i = 0; i < 16 && x + i < J.out_w; i++) {
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
        for (int ch = 0; ch < 3; ch++) {
          const int bidx = i * 3 + ch;
          o[ch] = (uint8_t)(w[r][bidx >> 2] >> (8 * (bidx & 3)));
        }
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
  const uint32_t groups = (job.out_w + 15) / 16;
  const uint32_t xblocks = (groups + 127) / 128, row_pairs = (job.out_h + 1) / 2;
  const unsigned grid = job.n_images * row_pairs * xblocks;
  if (job.fused) {
    if (job.full_range) color_stitch_kernel<true, true><<<grid, 128, 0, stream>>>(job, k, row_pairs, xblocks);
    else color_stitch_kernel<false, true><<<grid, 128, 0, stream>>>(job, k, row_pairs, xblocks);
  } else {
    if (job.full_range) color_stitch_kernel<true, false><<<grid, 128, 0, stream>>>(job, k, row_pairs, xblocks);
    else color_stitch_kernel<false, false><<<grid, 128, 0, stream>>>(job, k, row_pairs, xblocks);
  }
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
