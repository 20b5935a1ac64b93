// Sample adaptive offset (H.265 8.7.3) of eight horizontally adjacent samples: shared by the stand-alone SAO kernel
// (sao_kernel.cu) and the colour kernel that applies SAO on the fly (color_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace heic {
namespace dev {
namespace sao {

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

}  // namespace sao
}  // namespace dev
}  // namespace heic
