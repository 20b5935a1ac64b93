// Stage 4: deblocking filter (H.265 8.7.2), in place on the reconstruction arena.  Absent from the
// reference.  All-intra pictures: bS = 2 on every transform-block edge of the 8x8 grid (CU and PU edges
// are transform-block edges too), picture borders are not filtered, and HEIF grid tiles are separate
// pictures, so nothing crosses a tile.
//
// Edge-parallel formulation: the vertical-edge pass only touches columns [8k-4, 8k+4) around edge 8k and
// the horizontal-edge pass only rows [8j-4, 8j+4), so the "shifted" 8x8 cell [8k-4,8k+4) x [8j-4,8j+4) is
// closed under both passes: one thread loads it into registers, filters its vertical edge, then its
// horizontal edge on the result (the order 8.7.2 prescribes), and writes it back — one read and one write
// of every sample, no halo, no second pass over HBM.
#include <cuda_runtime.h>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

__device__ const uint8_t kBetaTable[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,
                                           8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32,
                                           34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
__device__ const uint8_t kTcTable[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  0,  0,
                                         1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2,  2,  2,  3,  3,  3,  3,  4,
                                         4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
__device__ const uint8_t kChromaQp[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};

__device__ __forceinline__ int clip3(int lo, int hi, int v) { return min(hi, max(lo, v)); }
__device__ __forceinline__ int clip8(int v) { return min(255, max(0, v)); }

__device__ __forceinline__ uint32_t interleave4(uint32_t v) {  // 4-bit coordinate -> even bits
  v &= 0x0fu;
  v = (v | (v << 2)) & 0x33u;
  v = (v | (v << 1)) & 0x55u;
  return v;
}

struct Pic {
  const uint32_t* tu_map;
  const uint8_t* qp_map;
  int w8, log2_ctb, wctb;
};

// Is the edge through luma sample (x, y) at coordinate `pos` (x for a vertical edge, y for a horizontal one; a multiple
// of 8) a transform-block edge?  Transform units of 4x4 and 8x8 always end on the 8-sample grid, and nothing is larger than
// 32x32, so the edge is NOT one only if a 16x16 or 32x32 unit covers the sample and `pos` is not a multiple of its size: two
// probes of tu_map (the 16- and the 32-aligned candidate origins) instead of four.  Both are always issued -- independent
// loads, no branch on the position: skipping them where `pos` makes them moot measured slower (14.5 vs 14.0 ms).
__device__ __forceinline__ bool is_tu_edge(const Pic& p, int x, int y, int pos) {
  const int ctb4 = 1 << (p.log2_ctb - 2);
  const int rx = x >> p.log2_ctb, ry = y >> p.log2_ctb;
  const uint32_t z = interleave4((uint32_t)(x >> 2) & (ctb4 - 1)) | (interleave4((uint32_t)(y >> 2) & (ctb4 - 1)) << 1);
  const uint32_t* tu = p.tu_map + (size_t)(ry * p.wctb + rx) * (ctb4 * ctb4);
  const uint32_t w3 = tu[z & ~63u], w2 = tu[z & ~15u];
  const bool in32 = (w3 & 7u) == (TU_ORIGIN | (3u << 1)) && (pos & 31) != 0;
  const bool in16 = (w2 & 7u) == (TU_ORIGIN | (2u << 1)) && (pos & 15) != 0;
  return !(in32 || in16);
}
__device__ __forceinline__ int qp_at(const Pic& p, int x, int y) { return p.qp_map[(y >> 3) * p.w8 + (x >> 3)]; }

// ---- the cell stays PACKED: rw[r][0] = samples 0..3 of row r (p3 p2 p1 p0 of the line across the vertical centre line),
// rw[r][1] = samples 4..7 (q0 q1 q2 q3).  Every quantity of 8.7.2.5.3 / 8.7.2.5.7 is a small-integer linear form of the eight
// samples of a line, so it is evaluated ON THE PACKED WORDS with two `dp4a.u32.s32` (unsigned samples x signed byte
// weights; the integer-multiplier pipe, which these kernels otherwise leave half idle) -- no unpacking.  And a filtered
// sample replaces the old one by ADDING its difference at the byte's position (three IMADs per word): every new sample is
// a clipped value in 0..255, so the sum of the per-byte differences never carries from one byte into the next -- no
// repacking either.  Line l of the segment the loop presents is row l (the loop swaps halves / transposes the cell so that
// all four segments of a cell take this form).
__device__ __forceinline__ int dp4(uint32_t samples, uint32_t weights, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(samples), "r"(weights), "r"(acc));
  return d;
}
// signed byte weights of (byte 0, byte 1, byte 2, byte 3) = (p3, p2, p1, p0) of the left word, (q0, q1, q2, q3) of the right one
__host__ __device__ constexpr uint32_t W(int a, int b, int c, int d) {
  return ((uint32_t)a & 0xffu) | (((uint32_t)b & 0xffu) << 8) | (((uint32_t)c & 0xffu) << 16) | (((uint32_t)d & 0xffu) << 24);
}
// P.w1 + Q.w2 + c
__device__ __forceinline__ int lin(uint32_t P, uint32_t Q, uint32_t wp, uint32_t wq, int c) { return dp4(Q, wq, dp4(P, wp, c)); }

// 8.7.2.5.3 decisions + 8.7.2.5.7 filters of the luma edge segment in rows 0..3.  Returns whether anything changed.
__device__ __forceinline__ bool filter_luma_segment(uint32_t (&rw)[8][2], int qp_sum, int beta_off2, int tc_off2) {
  const int qpl = (qp_sum + 1) >> 1;
  const int beta = kBetaTable[clip3(0, 51, qpl + beta_off2)];
  const int tc = kTcTable[clip3(0, 53, qpl + 2 + tc_off2)];
  const uint32_t P0 = rw[0][0], Q0 = rw[0][1], P3 = rw[3][0], Q3 = rw[3][1];
  const int dp0 = abs(dp4(P0, W(0, 1, -2, 1), 0)), dp3 = abs(dp4(P3, W(0, 1, -2, 1), 0));
  const int dq0 = abs(dp4(Q0, W(1, -2, 1, 0), 0)), dq3 = abs(dp4(Q3, W(1, -2, 1, 0), 0));
  const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
  if (dpq0 + dpq3 >= beta) return false;
  const int tc25 = (5 * tc + 1) >> 1;
  const bool s0 = 2 * dpq0 < (beta >> 2) && abs(dp4(P0, W(1, 0, 0, -1), 0)) + abs(dp4(Q0, W(1, 0, 0, -1), 0)) < (beta >> 3) &&
                  abs(lin(P0, Q0, W(0, 0, 0, 1), W(-1, 0, 0, 0), 0)) < tc25;
  const bool s3 = 2 * dpq3 < (beta >> 2) && abs(dp4(P3, W(1, 0, 0, -1), 0)) + abs(dp4(Q3, W(1, 0, 0, -1), 0)) < (beta >> 3) &&
                  abs(lin(P3, Q3, W(0, 0, 0, 1), W(-1, 0, 0, 0), 0)) < tc25;
  const bool strong = s0 && s3;
  const bool dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const uint32_t P = rw[l][0], Q = rw[l][1];
    if (strong) {
      // new - old of p0 p1 p2 q0 q1 q2: the old sample, times the divisor, is folded into the weights (the shift is exact on
      // it), then clipped to +-2tc
      const int t2 = 2 * tc;
      const int e_p0 = clip3(-t2, t2, lin(P, Q, W(0, 1, 2, 2 - 8), W(2, 1, 0, 0), 4) >> 3);
      const int e_p1 = clip3(-t2, t2, lin(P, Q, W(0, 1, 1 - 4, 1), W(1, 0, 0, 0), 2) >> 2);
      const int e_p2 = clip3(-t2, t2, lin(P, Q, W(2, 3 - 8, 1, 1), W(1, 0, 0, 0), 4) >> 3);
      const int e_q0 = clip3(-t2, t2, lin(P, Q, W(0, 0, 1, 2), W(2 - 8, 2, 1, 0), 4) >> 3);
      const int e_q1 = clip3(-t2, t2, lin(P, Q, W(0, 0, 0, 1), W(1, 1 - 4, 1, 0), 2) >> 2);
      const int e_q2 = clip3(-t2, t2, lin(P, Q, W(0, 0, 0, 1), W(1, 1, 3 - 8, 2), 4) >> 3);
      rw[l][0] = P + (uint32_t)e_p2 * 256u + (uint32_t)e_p1 * 65536u + (uint32_t)e_p0 * 16777216u;
      rw[l][1] = Q + (uint32_t)e_q0 + (uint32_t)e_q1 * 256u + (uint32_t)e_q2 * 65536u;
    } else {
      int delta = lin(P, Q, W(0, 0, 3, -9), W(9, -3, 0, 0), 8) >> 4;
      if (abs(delta) < tc * 10) {
        delta = clip3(-tc, tc, delta);
        const int p0 = dp4(P, W(0, 0, 0, 1), 0), q0 = dp4(Q, W(1, 0, 0, 0), 0);
        const int ep = clip8(p0 + delta) - p0, eq = clip8(q0 - delta) - q0;
        uint32_t addp = (uint32_t)ep * 16777216u, addq = (uint32_t)eq;
        if (dep) {
          const int p1 = dp4(P, W(0, 0, 1, 0), 0);
          const int e = clip3(-(tc >> 1), tc >> 1, (((dp4(P, W(0, 1, 0, 1), 1) >> 1) - p1 + delta) >> 1));
          addp += (uint32_t)(clip8(p1 + e) - p1) * 65536u;
        }
        if (deq) {
          const int q1 = dp4(Q, W(0, 1, 0, 0), 0);
          const int e = clip3(-(tc >> 1), tc >> 1, (((dp4(Q, W(1, 0, 1, 0), 1) >> 1) - q1 - delta) >> 1));
          addq += (uint32_t)(clip8(q1 + e) - q1) * 256u;
        }
        rw[l][0] = P + addp;
        rw[l][1] = Q + addq;
      }
    }
  }
  return true;
}

__device__ __forceinline__ bool filter_chroma_segment(uint32_t (&rw)[8][2], int qp_sum, int c_qp_off, int tc_off2) {
  const int qpi = ((qp_sum + 1) >> 1) + c_qp_off;  // cQpPicOffset: PPS offset only (8.7.2.5.5)
  const int qpc = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : kChromaQp[qpi - 30]);
  const int tc = kTcTable[clip3(0, 53, qpc + 2 + tc_off2)];
  if (!tc) return false;
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const uint32_t P = rw[l][0], Q = rw[l][1];
    const int delta = clip3(-tc, tc, lin(P, Q, W(0, 0, 1, -4), W(4, -1, 0, 0), 4) >> 3);
    const int p0 = dp4(P, W(0, 0, 0, 1), 0), q0 = dp4(Q, W(1, 0, 0, 0), 0);
    rw[l][0] = P + (uint32_t)(clip8(p0 + delta) - p0) * 16777216u;
    rw[l][1] = Q + (uint32_t)(clip8(q0 - delta) - q0);
  }
  return true;
}

// 8x8 byte transpose of the packed cell (rw[r][h]: bytes 4h..4h+3 of row r), one 4x4 quadrant at a time: 8 PRMT each.
__device__ __forceinline__ void transpose4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
  const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);  // a0 b0 a1 b1 | c0 d0 c1 d1
  const uint32_t t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);  // a2 b2 a3 b3 | c2 d2 c3 d3
  o0 = __byte_perm(t0, t1, 0x5410);
  o1 = __byte_perm(t0, t1, 0x7632);
  o2 = __byte_perm(t2, t3, 0x5410);
  o3 = __byte_perm(t2, t3, 0x7632);
}
__device__ __forceinline__ void transpose_cell(uint32_t (&rw)[8][2]) {
  uint32_t t[8][2];
  transpose4(rw[0][0], rw[1][0], rw[2][0], rw[3][0], t[0][0], t[1][0], t[2][0], t[3][0]);  // top-left stays
  transpose4(rw[4][0], rw[5][0], rw[6][0], rw[7][0], t[0][1], t[1][1], t[2][1], t[3][1]);  // bottom-left -> top-right
  transpose4(rw[0][1], rw[1][1], rw[2][1], rw[3][1], t[4][0], t[5][0], t[6][0], t[7][0]);  // top-right -> bottom-left
  transpose4(rw[4][1], rw[5][1], rw[6][1], rw[7][1], t[4][1], t[5][1], t[6][1], t[7][1]);
#pragma unroll
  for (int r = 0; r < 8; r++) rw[r][0] = t[r][0], rw[r][1] = t[r][1];
}

// One shifted cell of plane CIDX.  (k, j): cell indices; the cell covers plane samples
// [8k-4, 8k+4) x [8j-4, 8j+4) clipped to the picture.
//
// The four edge segments (vertical edge: rows 0..3, rows 4..7; then horizontal edge: columns 0..3, columns 4..7) run as
// ONE rolled loop over a single copy of the metadata probes and of the filter: each iteration filters the segment that
// lies in rows 0..3 across the vertical centre line, then swaps the cell's two halves; after the second iteration the cell
// is transposed, so the horizontal edge's segments take the same form, and transposed back at the end.  Four inlined
// copies of the filter (2.8 K instructions) made the kernel instruction-fetch bound (stall_no_instruction 3.4 per issue).
template <int CIDX>
__device__ __forceinline__ void deblock_cell(const Pic& pic, uint8_t* plane, int pitch, int pw, int ph, int k, int j,
                                             int beta_off2, int tc_off2, int c_qp_off) {
  constexpr int SUB = CIDX ? 1 : 0;
  const int x0 = 8 * k - 4, y0 = 8 * j - 4;
  const bool has_left = k > 0, has_right = 8 * k < pw, has_top = j > 0, has_bottom = 8 * j < ph;
  uint32_t rw[8][2];
  // rows y0 .. y0+7; the left half [x0, x0+4) exists when k > 0, the right half when 8k < pw
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const bool row_ok = r < 4 ? has_top : has_bottom;
    rw[r][0] = rw[r][1] = 0;
    if (row_ok) {
      const uint8_t* row = plane + (size_t)(y0 + r) * pitch;
      if (has_left) rw[r][0] = *reinterpret_cast<const uint32_t*>(row + x0);
      if (has_right) rw[r][1] = *reinterpret_cast<const uint32_t*>(row + x0 + 4);
    }
  }
  // vertical edge at x = 8k (luma: every 8 samples; chroma 4:2:0: every 8 chroma = 16 luma samples), segments of 4 rows;
  // horizontal edge at y = 8j, segments of 4 columns
  const int xe = (8 * k) << SUB, ye = (8 * j) << SUB;  // luma positions of the two edges
  // which of the four segments are transform-block edges, and their QPs: all metadata loads are issued before any filtering
  uint32_t seg_do = 0, seg_qp = 0;  // bit g; byte g: qp_p + qp_q
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const bool vert = g < 2;
    const int half = g & 1;
    bool ok;
    if (vert) ok = has_left && has_right && (half == 0 ? has_top : has_bottom);
    else ok = has_top && has_bottom && (half == 0 ? has_left : has_right);
    if (ok) {
      const int xl = vert ? xe : (x0 + 4 * half) << SUB, yl = vert ? (y0 + 4 * half) << SUB : ye;
      seg_qp |= (uint32_t)(qp_at(pic, xl, yl) + (vert ? qp_at(pic, xl - 1, yl) : qp_at(pic, xl, yl - 1))) << (8 * g);
      seg_do |= (is_tu_edge(pic, xl, yl, vert ? xl : yl) ? 1u : 0u) << g;
    }
  }
  if (!seg_do) return;
  bool changed = false;
  // vertical edge first, then the horizontal edge on the vertically filtered samples (the order 8.7.2 prescribes)
#pragma unroll 1
  for (int g = 0; g < 4; g++) {
    if ((seg_do >> g) & 1u) {
      const int qp_sum = (int)((seg_qp >> (8 * g)) & 0xffu);
      changed |= CIDX == 0 ? filter_luma_segment(rw, qp_sum, beta_off2, tc_off2) : filter_chroma_segment(rw, qp_sum, c_qp_off, tc_off2);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {  // the other segment of this edge moves into rows 0..3 (and back after it)
      uint32_t t0 = rw[r][0], t1 = rw[r][1];
      rw[r][0] = rw[r + 4][0], rw[r][1] = rw[r + 4][1];
      rw[r + 4][0] = t0, rw[r + 4][1] = t1;
    }
    if (g & 1) transpose_cell(rw);
  }
  if (!changed) return;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const bool row_ok = r < 4 ? has_top : has_bottom;
    if (!row_ok) continue;
    uint8_t* row = plane + (size_t)(y0 + r) * pitch;
    if (has_left) *reinterpret_cast<uint32_t*>(row + x0) = rw[r][0];
    if (has_right) *reinterpret_cast<uint32_t*>(row + x0 + 4) = rw[r][1];
  }
}

// One launch for luma (KIND 0) and one for the two chroma planes (KIND 1): each launch then runs a single filter body
// that fits the instruction cache.  grid: flat over (tile, block of cells of the tile), cells numbered row by row.
#ifndef HEIC_DEBLOCK_MIN_CTAS
#define HEIC_DEBLOCK_MIN_CTAS 12  // 40 registers: 11.7 -> 11.1 ms against 8 CTAs at 58 registers (16 at 32 registers: 11.2)
#endif
template <int KIND>
__global__ void __launch_bounds__(128, HEIC_DEBLOCK_MIN_CTAS) deblock_kernel(Arenas A, uint32_t blocks_per_tile) {
  const uint32_t tile = blockIdx.x / blocks_per_tile;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0 || tp->deblock_disabled) return;
  if (KIND == 1 && !pp->chroma) return;
  uint32_t g = (blockIdx.x % blocks_per_tile) * blockDim.x + threadIdx.x;
  const int pw = pp->w >> KIND, ph = pp->h >> KIND;
  const uint32_t cx = (uint32_t)((pw + 7) >> 3) + 1, cy = (uint32_t)((ph + 7) >> 3) + 1;
  int cidx = KIND;
  if (g >= cx * cy) {
    if (KIND == 0) return;
    g -= cx * cy;
    cidx = 2;
    if (g >= cx * cy) return;
  }
  const int k = (int)(g % cx), j = (int)(g / cx);
  if (8 * k - 4 >= pw || 8 * j - 4 >= ph) return;
  Pic pic;
  pic.tu_map = A.tu_map + tp->tu_off;
  pic.qp_map = A.qp_map + tp->map8_off;
  pic.w8 = pp->w8;
  pic.log2_ctb = pp->log2_ctb;
  pic.wctb = pp->wctb;
  const int beta_off2 = tp->beta_offset_div2 * 2, tc_off2 = tp->tc_offset_div2 * 2;
  deblock_cell<KIND>(pic, A.recon + tp->plane_off[cidx], KIND ? pp->pitch_c : pp->pitch_y, pw, ph, k, j, beta_off2, tc_off2,
                     cidx == 0 ? 0 : (cidx == 1 ? pp->pps_cb_qp_offset : pp->pps_cr_qp_offset));
}

}  // namespace

cudaError_t launch_deblock(const Arenas& A, uint32_t max_w, uint32_t max_h, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const uint32_t cells_y = (((max_w + 7) >> 3) + 1) * (((max_h + 7) >> 3) + 1);
  const uint32_t cells_c = 2 * ((((max_w >> 1) + 7) >> 3) + 1) * ((((max_h >> 1) + 7) >> 3) + 1);
  const uint32_t by = (cells_y + 127) / 128, bc = (cells_c + 127) / 128;
  deblock_kernel<0><<<A.n_tiles * by, 128, 0, stream>>>(A, by);
  deblock_kernel<1><<<A.n_tiles * bc, 128, 0, stream>>>(A, bc);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
