// Stage 6: YCbCr 4:2:0 -> RGB8 fused with HEIF grid stitching, conformance/canvas crop and (optionally)
// the irot rotation — one bandwidth-bound pass: 1.5 B read + 3 B written per output pixel.  Unrotated rows leave through
// shared memory as bulk copies (TMA, cp.async.bulk: UBLKCP in SASS); rotated output is transposed in shared memory.
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

// One 16-byte store row of the staging buffer = 48 bytes of RGB per thread: a CTA's 128 threads lay down 6144 contiguous
// bytes per canvas row, which go to HBM as ONE bulk copy (TMA, cp.async.bulk shared -> global).
constexpr int kColorThreads = 128;
constexpr int kStageVec = kColorThreads * 3;  // uint4 per staged row

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}

// A thread converts a 16 x 2 block of the canvas (two rows share one row of chroma): 32 + 16 bytes in, 96 bytes out.
// FULL: full-range input, where (256 Y + t) >> 8 == Y + (t >> 8): the chroma terms are computed once per 2x2 quad, and the
//   two pixels of a row that share them are clamped together by one VIADDMNMX.S16x2.RELU (max(min(Y + t, 255), 0) on both
//   16-bit halves), so a pixel costs half an instruction per channel plus the byte shuffles.
// FUSED: the planes are the deblocked reconstruction; CTBs with SAO on are read from the final arena (see convert_block).
// BULK: unrotated output whose rows are multiples of 16 bytes: the RGB of the CTA's strip is staged in shared memory and
//   stored by two bulk copies (one per row) instead of 3 x 16-byte stores per thread and row at a 48-byte stride.
// grid: flat over (image, pair of canvas rows, block of 16-pixel groups).
// RGB of the 16 x 2 block at canvas position (x, y) of `image` as 2 x 12 packed words (row r: w[r][0..11] = 48 bytes).
template <bool FULL, bool FUSED>
__device__ __forceinline__ void convert_block(const ColorJob& J, const Coeffs& K, uint32_t image, uint32_t x, uint32_t y, bool two_rows,
                                              uint32_t (&w)[2][12]) {
  const uint32_t tr = y / J.tile_h, ly = y - tr * J.tile_h;
  uint32_t yv[2][4], cbv[2], crv[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t xh = x + 8 * h;
    if (xh < J.out_w) {
      const uint32_t tc = xh / J.tile_w, lx = xh - tc * J.tile_w;
      const uint32_t ti = image * J.grid_cols * J.grid_rows + tr * J.grid_cols + tc;
      const uint8_t* t = J.planes + (size_t)ti * J.tile_stride;
      const uint8_t *ty = t, *tcb = t, *tcr = t;
      if (FUSED) {
        // the eight luma samples (both rows: ly is even) and the four chroma samples lie in one CTB; a component whose SAO
        // type is non-zero there has its final samples in the other arena (written by sao_sparse_kernel), every other one
        // is final as deblocked.  One 16-byte load of the CTB's parameter words decides (they are zero where SAO is off).
        const uint4 sp = *reinterpret_cast<const uint4*>(J.sao + (size_t)ti * J.sao_stride +
                                                         (size_t)((ly >> J.log2_ctb) * J.wctb + (lx >> J.log2_ctb)) * 4);
        const uint8_t* t2 = J.planes_sao + (size_t)ti * J.tile_stride;
        if (sp.x & 3u) ty = t2;
        if (sp.y & 3u) tcb = t2;
        if (sp.z & 3u) tcr = t2;
      }
      const uint2 a = *reinterpret_cast<const uint2*>(ty + (size_t)ly * J.pitch_y + lx);
      const uint2 b = two_rows ? *reinterpret_cast<const uint2*>(ty + (size_t)(ly + 1) * J.pitch_y + lx) : make_uint2(0u, 0u);
      cbv[h] = crv[h] = 0x80808080u;
      if (J.chroma) {
        const size_t co = (size_t)(ly >> 1) * J.pitch_c + (lx >> 1);
        cbv[h] = *reinterpret_cast<const uint32_t*>(tcb + J.cb_off + co);
        crv[h] = *reinterpret_cast<const uint32_t*>(tcr + J.cr_off + co);
      }
      yv[0][2 * h] = a.x, yv[0][2 * h + 1] = a.y, yv[1][2 * h] = b.x, yv[1][2 * h + 1] = b.y;
    } else {
      yv[0][2 * h] = yv[0][2 * h + 1] = yv[1][2 * h] = yv[1][2 * h + 1] = 0u;
      cbv[h] = crv[h] = 0x80808080u;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {  // groups of four pixels -> three output words per row
    if (FULL) {
      uint32_t rg[2][2], bb[2][2];  // [row][pixel pair]: R | G << 8 and B of the two pixels in the 16-bit halves
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int j = 2 * q + jj;  // chroma sample index 0..7
        const int c = (int)((cbv[j >> 2] >> (8 * (j & 3))) & 0xffu) - 128, d = (int)((crv[j >> 2] >> (8 * (j & 3))) & 0xffu) - 128;
        // the three chroma terms, each duplicated into both 16-bit halves
        const uint32_t r2 = __byte_perm((uint32_t)((K.rv * d + 128) >> 8), 0u, 0x1010);
        const uint32_t g2 = __byte_perm((uint32_t)((K.gu * c + K.gv * d + 128) >> 8), 0u, 0x1010);
        const uint32_t b2 = __byte_perm((uint32_t)((K.bu * c + 128) >> 8), 0u, 0x1010);
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const uint32_t y2 = __byte_perm(yv[r][q], 0u, jj ? 0x4342 : 0x4140);  // Y of the pair's two pixels, zero-extended
          const uint32_t R = __viaddmin_s16x2_relu(y2, r2, 0x00ff00ffu), G = __viaddmin_s16x2_relu(y2, g2, 0x00ff00ffu);
          bb[r][jj] = __viaddmin_s16x2_relu(y2, b2, 0x00ff00ffu);
          rg[r][jj] = __byte_perm(R, G, 0x6240);  // R0 G0 R1 G1
        }
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        w[r][3 * q + 0] = __byte_perm(rg[r][0], bb[r][0], 0x2410);                                     // R0 G0 B0 R1
        w[r][3 * q + 1] = __byte_perm(__byte_perm(rg[r][0], bb[r][0], 0x0063), rg[r][1], 0x5410);     // G1 B1 R2 G2
        w[r][3 * q + 2] = __byte_perm(bb[r][1], rg[r][1], 0x2760);                                     // B2 R3 G3 B3
      }
    } else {
      uint32_t px[2][4][3];
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        const int j = 2 * q + jj;
        const int c = (int)((cbv[j >> 2] >> (8 * (j & 3))) & 0xffu) - 128, d = (int)((crv[j >> 2] >> (8 * (j & 3))) & 0xffu) - 128;
        const int r_add = K.rv * d + 128, g_add = K.gu * c + K.gv * d + 128, b_add = K.bu * c + 128;
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int p = 2 * jj + i;  // pixel within the group
            const int yy = K.y_mul * ((int)((yv[r][q] >> (8 * p)) & 0xffu) - K.y_sub);
            px[r][p][0] = clip8((yy + r_add) >> 8);
            px[r][p][1] = clip8((yy + g_add) >> 8);
            px[r][p][2] = clip8((yy + b_add) >> 8);
          }
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        w[r][3 * q + 0] = pack4(px[r][0][0], px[r][0][1], px[r][0][2], px[r][1][0]);
        w[r][3 * q + 1] = pack4(px[r][1][1], px[r][1][2], px[r][2][0], px[r][2][1]);
        w[r][3 * q + 2] = pack4(px[r][2][2], px[r][3][0], px[r][3][1], px[r][3][2]);
      }
    }
  }
}

template <bool FULL, bool FUSED, bool BULK>
__global__ void __launch_bounds__(kColorThreads) color_stitch_kernel(ColorJob J, Coeffs K, uint32_t row_pairs, uint32_t xblocks) {
  __shared__ __align__(128) uint4 stage[BULK ? 2 * kStageVec : 1];
  const uint32_t per_image = row_pairs * xblocks;
  const uint32_t image = blockIdx.x / per_image, rem = blockIdx.x % per_image;
  const uint32_t y = (rem / xblocks) * 2;
  const uint32_t xb = (rem % xblocks) * kColorThreads * 16;
  const uint32_t x = xb + threadIdx.x * 16;
  const bool in_range = x < J.out_w;  // y < out_h by construction of the grid
  if (!BULK && !in_range) return;
  // tile widths are multiples of 8: the first and the second 8 pixels may lie in different tiles
  const bool two_rows = y + 1 < J.out_h;  // tile heights are even, so row y + 1 is in the same tile
  uint32_t w[2][12];  // packed RGB of the two rows
  if (in_range) convert_block<FULL, FUSED>(J, K, image, x, y, two_rows, w);
  uint8_t* img = J.rgb + (size_t)image * J.image_stride;
  if (BULK) {
    if (in_range) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        uint4* s = stage + r * kStageVec + threadIdx.x * 3;  // 16-byte stores at a 48-byte stride: conflict-free
        s[0] = make_uint4(w[r][0], w[r][1], w[r][2], w[r][3]);
        s[1] = make_uint4(w[r][4], w[r][5], w[r][6], w[r][7]);
        s[2] = make_uint4(w[r][8], w[r][9], w[r][10], w[r][11]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the writes above, visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = min(J.out_w - xb, (uint32_t)kColorThreads * 16u) * 3u;  // a multiple of 16 (out_w is one of 16)
      uint8_t* o = img + (size_t)y * J.pitch + (size_t)xb * 3;
      bulk_store(o, stage, bytes);
      if (two_rows) bulk_store(o + J.pitch, stage + kStageVec, bytes);
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // the staging buffer lives until it has been read
    }
  } else {
#pragma unroll
  for (int r = 0; r < 2; r++) {
    if (r == 1 && !two_rows) break;
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
  }
  }
}

// irot (ISO/IEC 23008-12 6.5.10: anti-clockwise quarter turns) applied on the way out.  A CTA converts a 64 x 64 tile of the
// canvas and lays it down ROTATED in shared memory, so that the rows of the rotated picture leave as contiguous runs of
// 16-byte stores (a per-pixel scatter to HBM wrote 3 bytes per store).  canvas (sx, sy) -> rotated (dx, dy):
//   1: (sy, W - 1 - sx)    2: (W - 1 - sx, H - 1 - sy)    3: (H - 1 - sy, sx)
constexpr int kRotTile = 64;
constexpr int kRotRowBytes = kRotTile * 3 + 16;  // padded row of the rotated tile (a multiple of 16)

template <bool FULL, bool FUSED>
__global__ void __launch_bounds__(kColorThreads) color_rotate_kernel(ColorJob J, Coeffs K, uint32_t tiles_x, uint32_t tiles_y) {
  __shared__ __align__(16) uint8_t tile[kRotTile * kRotRowBytes];
  const uint32_t per_image = tiles_x * tiles_y;
  const uint32_t image = blockIdx.x / per_image, rem = blockIdx.x % per_image;
  const uint32_t x0 = (rem % tiles_x) * kRotTile, y0 = (rem / tiles_x) * kRotTile;
  const uint32_t w_v = min((uint32_t)kRotTile, J.out_w - x0), h_v = min((uint32_t)kRotTile, J.out_h - y0);  // valid part of the tile
  // a warp takes a 16-pixel column strip of the tile, a lane two of its rows
  const uint32_t lx = (threadIdx.x >> 5) * 16, ly = (threadIdx.x & 31) * 2;
  const uint32_t rot = J.rotation;
  if (lx < w_v && ly < h_v) {
    const bool two_rows = ly + 1 < h_v;
    uint32_t w[2][12];
    convert_block<FULL, FUSED>(J, K, image, x0 + lx, y0 + ly, y0 + ly + 1 < J.out_h, w);
    // byte b (0..47) of row r's 48 bytes of RGB
#define HEIC_RGB_BYTE(r, b) ((w[r][(b) >> 2] >> (8 * ((b) & 3))) & 0xffu)
    // two such bytes as one 16-bit value (low byte first): one PRMT
#define HEIC_RGB_PAIR(ra, ba, rb, bb) __byte_perm(w[ra][(ba) >> 2], w[rb][(bb) >> 2], ((ba) & 3) | ((4 + ((bb) & 3)) << 4))
    if (two_rows && lx + 16 <= w_v && rot != 2 && !(h_v & 1u)) {
      // quarter turns: the two rows of a column are two neighbouring pixels of one rotated row -- six contiguous bytes at an
      // even address, written as three 16-bit stores (rot 3: the lower canvas row comes first, rot 1: the upper one)
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t sx = lx + i;
        uint8_t* o = rot == 1 ? tile + (w_v - 1 - sx) * kRotRowBytes + ly * 3 : tile + sx * kRotRowBytes + (h_v - 2 - ly) * 3;
        uint16_t* o16 = reinterpret_cast<uint16_t*>(o);
        if (rot == 1) {
          o16[0] = (uint16_t)HEIC_RGB_PAIR(0, 3 * i, 0, 3 * i + 1);
          o16[1] = (uint16_t)HEIC_RGB_PAIR(0, 3 * i + 2, 1, 3 * i);
          o16[2] = (uint16_t)HEIC_RGB_PAIR(1, 3 * i + 1, 1, 3 * i + 2);
        } else {
          o16[0] = (uint16_t)HEIC_RGB_PAIR(1, 3 * i, 1, 3 * i + 1);
          o16[1] = (uint16_t)HEIC_RGB_PAIR(1, 3 * i + 2, 0, 3 * i);
          o16[2] = (uint16_t)HEIC_RGB_PAIR(0, 3 * i + 1, 0, 3 * i + 2);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (r == 1 && !two_rows) break;
        const uint32_t sy = ly + r;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint32_t sx = lx + i;
          if (sx < w_v) {
            uint32_t row, col;  // position inside the rotated tile
            if (rot == 1) row = w_v - 1 - sx, col = sy;
            else if (rot == 2) row = h_v - 1 - sy, col = w_v - 1 - sx;
            else row = sx, col = h_v - 1 - sy;
            uint8_t* o = tile + row * kRotRowBytes + col * 3;
            o[0] = (uint8_t)HEIC_RGB_BYTE(r, 3 * i);
            o[1] = (uint8_t)HEIC_RGB_BYTE(r, 3 * i + 1);
            o[2] = (uint8_t)HEIC_RGB_BYTE(r, 3 * i + 2);
          }
        }
      }
    }
#undef HEIC_RGB_BYTE
#undef HEIC_RGB_PAIR
  }
  __syncthreads();
  // rotated tile -> HBM
  uint32_t n_rows, n_cols, dx0, dy0;
  if (rot == 1) n_rows = w_v, n_cols = h_v, dx0 = y0, dy0 = J.out_w - x0 - w_v;
  else if (rot == 2) n_rows = h_v, n_cols = w_v, dx0 = J.out_w - x0 - w_v, dy0 = J.out_h - y0 - h_v;
  else n_rows = w_v, n_cols = h_v, dx0 = J.out_h - y0 - h_v, dy0 = x0;
  uint8_t* dst = J.rgb + (size_t)image * J.image_stride + (size_t)dy0 * J.pitch + (size_t)dx0 * 3;
  const uint32_t row_bytes = n_cols * 3;
  if (((row_bytes | (uint32_t)J.pitch | (uint32_t)((uintptr_t)dst)) & 15u) == 0) {
    const uint32_t chunks = row_bytes >> 4;
    for (uint32_t q = threadIdx.x; q < n_rows * chunks; q += kColorThreads) {
      const uint32_t row = q / chunks, c = q - row * chunks;
      *reinterpret_cast<uint4*>(dst + (size_t)row * J.pitch + c * 16) = *reinterpret_cast<const uint4*>(tile + row * kRotRowBytes + c * 16);
    }
  } else {
    for (uint32_t q = threadIdx.x; q < n_rows * row_bytes; q += kColorThreads) {
      const uint32_t row = q / row_bytes, c = q - row * row_bytes;
      dst[(size_t)row * J.pitch + c] = tile[row * kRotRowBytes + c];
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
  const uint32_t xblocks = (groups + kColorThreads - 1) / kColorThreads, row_pairs = (job.out_h + 1) / 2;
  const unsigned grid = job.n_images * row_pairs * xblocks;
  // bulk (TMA) stores need 16-byte aligned rows whose length is a multiple of 16 bytes
  const bool bulk = job.rotation == 0 && job.out_w % 16 == 0 && job.pitch % 16 == 0 && job.image_stride % 16 == 0 &&
                    ((uintptr_t)job.rgb & 15u) == 0;
  if (job.rotation != 0) {
    const uint32_t tiles_x = (job.out_w + kRotTile - 1) / kRotTile, tiles_y = (job.out_h + kRotTile - 1) / kRotTile;
    const unsigned rgrid = job.n_images * tiles_x * tiles_y;
    if (job.fused) {
      if (job.full_range) color_rotate_kernel<true, true><<<rgrid, kColorThreads, 0, stream>>>(job, k, tiles_x, tiles_y);
      else color_rotate_kernel<false, true><<<rgrid, kColorThreads, 0, stream>>>(job, k, tiles_x, tiles_y);
    } else {
      if (job.full_range) color_rotate_kernel<true, false><<<rgrid, kColorThreads, 0, stream>>>(job, k, tiles_x, tiles_y);
      else color_rotate_kernel<false, false><<<rgrid, kColorThreads, 0, stream>>>(job, k, tiles_x, tiles_y);
    }
    return cudaGetLastError();
  }
#define HEIC_COLOR_LAUNCH(FULL, FUSED)                                                                                  \
  do {                                                                                                                  \
    if (bulk) color_stitch_kernel<FULL, FUSED, true><<<grid, kColorThreads, 0, stream>>>(job, k, row_pairs, xblocks);   \
    else color_stitch_kernel<FULL, FUSED, false><<<grid, kColorThreads, 0, stream>>>(job, k, row_pairs, xblocks);       \
  } while (0)
  if (job.fused) {
    if (job.full_range) HEIC_COLOR_LAUNCH(true, true);
    else HEIC_COLOR_LAUNCH(false, true);
  } else {
    if (job.full_range) HEIC_COLOR_LAUNCH(true, false);
    else HEIC_COLOR_LAUNCH(false, false);
  }
#undef HEIC_COLOR_LAUNCH
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
