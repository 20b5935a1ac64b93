// Stage 2: scaling (H.265 8.6.4.2 with the 7.4.5 scaling factors) + inverse DST-4 / DCT-4..32 (8.6.4.2),
// in place on the coefficient arena.  The reference has none of this (slice.rs:253-255 is todo!()).
//
// Small integer butterflies.  A first pass over tu_map appends every coded block (cbf = 1) to the dense list of its
// transform size; then one launch per size works through its list, so every warp step is full of coded blocks and every
// CTA of a launch runs the same straight-line butterfly (a single kernel with all sizes was instruction-fetch bound, and
// sweeping tu_map in place left most lanes of a step without a block):
//   n >= 8: 32/n blocks per warp step; a lane runs one n-point 1-D transform per pass (columns, then rows) as an
//           even/odd partial butterfly in registers; the passes are joined through a padded (conflict-free)
//           shared-memory tile and the residual goes back to HBM in lane-contiguous words;
//   n == 4: one lane per block, both passes in registers, 32-byte loads/stores.
// Only coded blocks are read or written, so DRAM traffic is 2 B in + 2 B out per coded sample (+ 8 B per block of list).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

__host__ __device__ constexpr int dct_quarter(int m) {
  constexpr int q[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                         61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
  return q[m];
}
// transMatrix coefficient of the 32-point DCT (8.6.4.2): row k, column n = C[(k * (2n + 1)) mod 128] with
// the cosine's symmetries folded.
__host__ __device__ constexpr int dct32(int k, int n) {
  int m = (k * (2 * n + 1)) & 127;
  return m <= 32 ? dct_quarter(m) : m <= 64 ? -dct_quarter(64 - m) : m < 96 ? -dct_quarter(m - 64) : dct_quarter(128 - m);
}

// y[k] = sum_j dct_N[j][k] * x[j] for j < NZ (inputs at j >= NZ are known to be zero).
template <int N, int NZ>
struct Idct {
  static __device__ __forceinline__ void run(const int (&x)[N], int (&y)[N]) {
    constexpr int H = N / 2, HZ = (NZ + 1) / 2;
    int xe[H], e[H];
#pragma unroll
    for (int j = 0; j < H; j++) xe[j] = x[2 * j];
    Idct<H, (HZ < 1 ? 1 : HZ)>::run(xe, e);
#pragma unroll
    for (int k = 0; k < H; k++) {
      int o = 0;
#pragma unroll
      for (int j = 1; j < N; j += 2)
        if (j < NZ) o += dct32(j * (32 / N), k) * x[j];
      y[k] = e[k] + o;
      y[N - 1 - k] = e[k] - o;
    }
  }
};
template <int NZ>
struct Idct<2, NZ> {
  static __device__ __forceinline__ void run(const int (&x)[2], int (&y)[2]) {
    int a = 64 * x[0], b = NZ > 1 ? 64 * x[1] : 0;
    y[0] = a + b;
    y[1] = a - b;
  }
};

__device__ __forceinline__ void dst4(const int (&x)[4], int (&y)[4]) {
  y[0] = 29 * x[0] + 74 * x[1] + 84 * x[2] + 55 * x[3];
  y[1] = 55 * x[0] + 74 * x[1] - 29 * x[2] - 84 * x[3];
  y[2] = 74 * x[0] - 74 * x[2] + 74 * x[3];
  y[3] = 84 * x[0] - 74 * x[1] + 55 * x[2] - 29 * x[3];
}

__device__ __forceinline__ int clip16(int v) { return min(32767, max(-32768, v)); }

template <int N, bool DST>
__device__ __forceinline__ void transform_1d(const int (&x)[N], int (&y)[N], int nz) {
  if (DST) {
    int xx[4] = {x[0], x[1], x[2], x[3]}, yy[4];
    dst4(xx, yy);
#pragma unroll
    for (int i = 0; i < 4; i++) y[i] = yy[i];
  } else if (N >= 16 && nz <= N / 4) {
    Idct<N, (N / 4 < 1 ? 1 : N / 4)>::run(x, y);
  } else if (N >= 8 && nz <= N / 2) {
    Idct<N, N / 2>::run(x, y);
  } else {
    Idct<N, N>::run(x, y);
  }
}

__device__ const uint8_t kLevelScale[6] = {40, 45, 51, 57, 64, 72};
__device__ const uint8_t kChromaQp[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};  // qPi 30..43

constexpr int kWarpsPerCta = 8;
constexpr int kTmpPerWarp = 32 * 34;  // int16 elements: one 32x32 block with padded rows

// ---- list construction ---------------------------------------------------------------------------------------
// A CTA classifies kListPerCta consecutive tu_map entries of one tile, sixteen per thread.  The per-class item counts are
// packed in one 64-bit word (12 bits per class) so a single warp scan places every thread's items; a warp reserves its
// space in the CTA with shared-memory atomics and the CTA reserves its space in every list with one global atomic per
// class.  Item order within a list is not deterministic; the blocks are independent, so the results are.
constexpr int kListThreads = 256;
constexpr int kListIter = 4;                              // 16-byte loads per thread
constexpr int kListPerCta = 4 * kListThreads * kListIter;  // tu_map entries per CTA (16 KB)

// classes of the (up to) three coded blocks of a tu_map word, 3 bits each (7: none)
__device__ __forceinline__ uint32_t classify(uint32_t w, bool chroma) {
  uint32_t cy = 7, cb = 7, cr = 7;
  if (w & TU_ORIGIN) {
    const int lg = (int)tu_log2(w);
    if (w & TU_CBF_Y) cy = 5 - lg;  // LIST_32 .. LIST_4Y
    // chroma rides on the luma TU of twice its size, or on blkIdx 3 of a split 8x8
    if (chroma && (lg > 2 || (w & TU_HAS_CHROMA))) {
      const uint32_t cc = lg == 5 ? LIST_16 : lg == 4 ? LIST_8 : LIST_4C;
      if (w & TU_CBF_CB) cb = cc;
      if (w & TU_CBF_CR) cr = cc;
    }
  }
  return cy | (cb << 3) | (cr << 6);
}

#ifndef HEIC_LIST_MIN_CTAS
#define HEIC_LIST_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(kListThreads, HEIC_LIST_MIN_CTAS) tu_list_kernel(Arenas A, uint32_t blocks_per_tile) {
  __shared__ uint32_t cta_cnt[LIST_CLASSES], cta_base[LIST_CLASSES], warp_base[kListThreads / 32][LIST_CLASSES];
  const uint32_t tile = blockIdx.x / blocks_per_tile;
  const uint32_t first = (blockIdx.x % blocks_per_tile) * kListPerCta;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (A.status[tile].code != 0 || first >= (uint32_t)pp->n_tu) return;  // uniform over the CTA
  if (threadIdx.x < LIST_CLASSES) cta_cnt[threadIdx.x] = 0;
  __syncthreads();
  const bool chroma = pp->chroma != 0;
  // kListIter coalesced 16-byte loads per thread (n_tu and tu_off are multiples of 16)
  uint32_t cls[kListIter][4], w_all[kListIter][4];
  unsigned long long mine = 0;  // items of this thread per class, 12 bits each
#pragma unroll
  for (int it = 0; it < kListIter; it++) {
    const uint32_t i0 = first + (uint32_t)it * (4 * kListThreads) + threadIdx.x * 4;
    uint4 wv = make_uint4(0u, 0u, 0u, 0u);
    if (i0 < (uint32_t)pp->n_tu) wv = *reinterpret_cast<const uint4*>(A.tu_map + tp->tu_off + i0);
    const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      w_all[it][e] = w[e];
      cls[it][e] = 0x1ffu;
      if (w[e]) {
        cls[it][e] = classify(w[e], chroma);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const uint32_t k = (cls[it][e] >> (3 * c)) & 7u;
          if (k < LIST_CLASSES) mine += 1ull << (12 * k);
        }
      }
    }
  }
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
  if (lane < LIST_CLASSES) {
    const uint32_t n = (uint32_t)(total >> (12 * lane)) & 0xfffu;
    warp_base[warp][lane] = n ? atomicAdd(cta_cnt + lane, n) : 0u;
  }
  __syncthreads();
  if (threadIdx.x < LIST_CLASSES)
    cta_base[threadIdx.x] = A.list_off[threadIdx.x] +
                            (cta_cnt[threadIdx.x] ? atomicAdd(A.list_count + threadIdx.x * LIST_COUNT_STRIDE, cta_cnt[threadIdx.x]) : 0u);
  __syncthreads();
  if (!mine) return;
  unsigned long long run = incl - mine;
#pragma unroll
  for (int it = 0; it < kListIter; it++) {
    const uint32_t i0 = first + (uint32_t)it * (4 * kListThreads) + threadIdx.x * 4;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      if (cls[it][e] == 0x1ffu) continue;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const uint32_t k = (cls[it][e] >> (3 * c)) & 7u;
        if (k < LIST_CLASSES) {
          // the block's QpY and transform_skip_flag ride in the item, so the transform kernels never touch tu_map again
          // (for a 4x4 block that read was a 32-byte sector per 64 bytes of payload)
          const uint32_t we = w_all[it][e];
          uint2_t v;
          v.x = tile | (tu_qp(we) << 24) | (tu_tskip(we, c) << 30);
          v.y = (i0 + e) | ((uint32_t)c << 30);
          A.tu_list[(size_t)cta_base[k] + warp_base[warp][k] + ((uint32_t)(run >> (12 * k)) & 0xfffu)] = v;
          run += 1ull << (12 * k);
        }
      }
    }
  }
}

// A coded block named by a list item.
struct Item {
  const PicParams* pp;
  const TileParams* tp;
  int16_t* blk;
  int qp, tskip;
  int cidx;
};
__device__ __forceinline__ Item fetch_item(const Arenas& A, uint2_t it) {
  Item r;
  r.tp = A.tiles + (it.x & 0xffffffu);
  r.pp = A.pics + r.tp->pic;
  const uint32_t entry = it.y & 0x3fffffffu;
  r.cidx = (int)(it.y >> 30);
  r.qp = (int)((it.x >> 24) & 63u);
  r.tskip = (int)((it.x >> 30) & 1u);
  // luma: 16 coefficients per tu_map entry; chroma: 4, at the entry of the 8x8 luma area the block belongs to
  r.blk = A.coeff + r.tp->coeff_off[r.cidx] + (r.cidx ? (size_t)(entry & ~3u) * 4 : (size_t)entry * 16);
  return r;
}
__device__ __forceinline__ int item_scale(const Item& t) {  // levelScale[qP % 6] << (qP / 6), 8.6.2 / 8.6.4.2
  int qp = t.qp;
  if (t.cidx) {
    int qpi = qp + (t.cidx == 1 ? t.pp->pps_cb_qp_offset + t.tp->slice_cb_qp_offset : t.pp->pps_cr_qp_offset + t.tp->slice_cr_qp_offset);
    qpi = min(57, max(0, qpi));
    qp = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : kChromaQp[qpi - 30]);
  }
  return (int)kLevelScale[qp % 6] << (qp / 6);
}

// ---- n >= 8: one launch per size over its list ------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(kWarpsPerCta * 32) transform_kernel(Arenas A, int cls) {
  constexpr int K = 32 / N;  // blocks per warp step
  constexpr int LOG2 = N == 8 ? 3 : N == 16 ? 4 : 5;
  constexpr int S = N + 2;   // padded row stride (int16) of the transpose tile
  __shared__ __align__(16) int16_t tmp_all[kWarpsPerCta][kTmpPerWarp];
  __shared__ int16_t* blk_of[kWarpsPerCta][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = lane / N, col = lane % N;
  const uint32_t count = A.list_count[cls * LIST_COUNT_STRIDE];
  const uint2_t* list = A.tu_list + A.list_off[cls];
  int16_t* tmp = tmp_all[warp];
  int16_t* t = tmp + k * (N * S);
  const uint32_t steps = (count + K - 1) / K;
  for (uint32_t step = blockIdx.x * kWarpsPerCta + warp; step < steps; step += gridDim.x * kWarpsPerCta) {
    const uint32_t idx = step * K + k;
    const bool active = idx < count;
    uint2_t it;
    it.x = it.y = 0;
    if (active) it = list[idx];
    const Item item = fetch_item(A, it);  // item 0 of tile 0 for the idle lanes of the last step: read, never written
    int16_t* blk = item.blk;
    if (col == 0) blk_of[warp][k] = blk;
    const bool tskip = active && item.tskip;
    // ---- DC-only blocks (every block of this warp step has at most its DC coefficient; most large blocks of smooth
    // pictures): both 1-D passes of a lone DC term are a multiplication by 64, so every sample of the block is the same
    // value -- no butterflies, no transpose tile, just the fill.  Decided on the raw levels, warp-uniformly.
    int x[N], y[N];
#pragma unroll
    for (int j = 0; j < N; j++) x[j] = active ? (int)blk[j * N + col] : 0;  // raw levels of this lane's column
    {
      int any_other = 0;
#pragma unroll
      for (int j = 0; j < N; j++) any_other |= (j == 0 && col == 0) ? 0 : x[j];
      if (!__any_sync(0xffffffffu, any_other != 0 || tskip)) {
        int v = 0;
        if (active && col == 0) {
          const int lvl = x[0];
          const int scale = item_scale(item);
          int mf = 16;
          if (item.pp->scaling_enabled) {
            const ScalingSet* sc = A.scaling + item.pp->scaling_set;
            mf = N == 8 ? sc->f8[item.cidx][0] : N == 16 ? sc->f16[item.cidx][0] : sc->f32[item.cidx][0];
          }
          constexpr int BD_SHIFT = LOG2 + 3;
          long long p = ((long long)lvl * (mf * scale) + (1ll << (BD_SHIFT - 1))) >> BD_SHIFT;
          const int dc = (int)min(32767ll, max(-32768ll, p));
          const int t1 = clip16((64 * dc + 64) >> 7);
          v = clip16((64 * t1 + 2048) >> 12);
        }
        v = __shfl_sync(0xffffffffu, v, k * N);  // the block's lane 0
        const uint32_t vv = ((uint32_t)v & 0xffffu) * 0x10001u;
        constexpr int PAIRS = N * N / 2;
        const uint32_t n_active = min((uint32_t)K, count - step * K);
#pragma unroll
        for (int g = 0; g < K; g++) {
          if ((uint32_t)g >= n_active) break;  // warp-uniform
          const uint32_t vg = __shfl_sync(0xffffffffu, vv, g * N);
          uint4* dst = reinterpret_cast<uint4*>(__shfl_sync(0xffffffffu, (unsigned long long)blk, g * N));
#pragma unroll 4
          for (int q = lane; q < PAIRS / 4; q += 32) dst[q] = make_uint4(vg, vg, vg, vg);
        }
        continue;
      }
    }
    // The two 1-D passes run as one loop; for N == 32 it is not unrolled, so the three butterfly variants exist once
    // instead of twice (3.5 K instructions did not fit the instruction cache: stall_no_instruction 5.7 per issue).
    int nz2 = 0;
    constexpr int UNROLL = N == 32 ? 1 : 2;
#pragma unroll UNROLL
    for (int pass = 0; pass < 2; pass++) {
      int nz;
      if (pass == 0) {
        // ---- scaling + first (column) pass -----------------------------------------------------
        int nz_rows = 0;
        // raw levels first: only the rows up to the last non-zero one (warp-uniform bound) are scaled afterwards
#pragma unroll
        for (int j = 0; j < N; j++) nz_rows = x[j] ? j + 1 : nz_rows;
        // extents of the non-zero coefficients, uniform over the warp so the butterfly variant is too
        int nz1 = nz_rows;
#pragma unroll
        for (int o = 16; o; o >>= 1) nz1 = max(nz1, __shfl_xor_sync(0xffffffffu, nz1, o));
        const uint32_t col_mask = __ballot_sync(0xffffffffu, nz_rows > 0);
        // highest non-zero column index + 1 within any block of the step
#pragma unroll
        for (int g = 0; g < K; g++) {
          uint32_t mg = (col_mask >> (g * N)) & (N == 32 ? 0xffffffffu : ((1u << N) - 1u));
          nz2 = max(nz2, 32 - __clz(mg));
        }
        if (active) {
          const int scale = item_scale(item);
          const uint8_t* m = nullptr;
          if (item.pp->scaling_enabled && !tskip) {
            const ScalingSet* sc = A.scaling + item.pp->scaling_set;
            m = N == 8 ? sc->f8[item.cidx] : N == 16 ? sc->f16[item.cidx] : sc->f32[item.cidx];
          }
          constexpr int BD_SHIFT = LOG2 + 3;  // BitDepth + log2(nTbS) - 5, 8-bit
#pragma unroll
          for (int j = 0; j < N; j++) {
            if (j < nz1) {  // warp-uniform
              const int lvl = x[j];
              const int ms = (m ? (int)m[j * N + col] : 16) * scale;  // <= 255 * (72 << 8) < 2^23
              int v;
              if ((unsigned)(lvl + 255) <= 510u) {  // |level| <= 255: the product fits 32 bits (the common case by far)
                v = (lvl * ms + (1 << (BD_SHIFT - 1))) >> BD_SHIFT;
              } else {
                long long p = (long long)lvl * ms + (1ll << (BD_SHIFT - 1));
                p >>= BD_SHIFT;
                v = (int)min(32767ll, max(-32768ll, p));
              }
              x[j] = min(32767, max(-32768, v));
            }
          }
        }
        nz = nz1;
      } else {
        // ---- second (row) pass: lane `col` now owns row `col` of its block ---------------------------
        const int16_t* row = t + col * S;
#pragma unroll
        for (int j = 0; j < N; j += 2) {
          uint32_t two = *reinterpret_cast<const uint32_t*>(row + j);
          x[j] = (int)(int16_t)(two & 0xffffu);
          x[j + 1] = (int)(int16_t)(two >> 16);
        }
        nz = nz2;
      }
      if (tskip) {
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = pass ? ((x[i] << 7) + 2048) >> 12 : x[i];  // rotation in the second pass
      } else {
        transform_1d<N, false>(x, y, nz);
        const int add = pass ? 2048 : 64, sh = pass ? 12 : 7;
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = clip16((y[i] + add) >> sh);
      }
      __syncwarp();
      if (pass == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) t[i * S + col] = (int16_t)y[i];
      } else {
        int16_t* row = t + col * S;
#pragma unroll
        for (int j = 0; j < N; j += 2)
          *reinterpret_cast<uint32_t*>(row + j) = ((uint32_t)y[j] & 0xffffu) | ((uint32_t)y[j + 1] << 16);
      }
      __syncwarp();
    }
    // ---- coalesced store of the residual blocks (pairs of int16) ----------------------------------
    {
      constexpr int PAIRS = N * N / 2;  // per block
      const uint32_t n_active = min((uint32_t)K, count - step * K);
#pragma unroll
      for (int g = 0; g < K; g++) {
        if ((uint32_t)g >= n_active) break;  // warp-uniform
        uint32_t* dst = reinterpret_cast<uint32_t*>(blk_of[warp][g]);
        const int16_t* src = tmp + g * (N * S);
#pragma unroll 4
        for (int q = lane; q < PAIRS; q += 32) {
          const int r = q / (N / 2), cp = q % (N / 2);
          dst[q] = *reinterpret_cast<const uint32_t*>(src + r * S + 2 * cp);
        }
      }
    }
    __syncwarp();
  }
}

// ---- n = 16, 32 on the tensor cores: the block as two integer matrix products ---------------------------------------
// T = M^T X (columns, + 64 >> 7, clip) and R = T M (rows, + 2048 >> 12, clip) are dense contractions over 16 / 32 terms with
// an 8-bit constant matrix.  A 16-bit operand is split as 256 * hi (signed byte) + lo (unsigned byte), so each product is two
// `mma.sync.m16n8k{16,32}` with s8 x s8 and s8 x u8 operands and exact s32 accumulation (|sum| <= 32 * 90 * 32768 < 2^31): the
// same integers as the butterflies, hence the same clipped results.
//  * The contraction index may be permuted freely as long as both operands agree.  `ldmatrix.trans` hands a thread the
//    coefficients X[8q + 2t][c], X[8q + 2t + 1][c] (q = 0..3) of its column c; their high / low bytes gathered with one
//    PRMT each are the B fragment of the column pass for the order j = {2t, 2t+1, 8+2t, 9+2t | 16+2t, 17+2t, 24+2t, 25+2t}.
//  * In that same order the accumulator fragments of the column pass (thread: T[r][8n + 2t], T[r][8n + 2t + 1]) ARE the A
//    fragments of the row pass -- the two passes are joined in registers, no shared-memory transpose, no shuffles.
//  * One set of eight constant registers per thread (M in the permuted order) serves as A of the first pass and B of the
//    second.
// One block per warp step; levels are loaded and scaled 16 bytes at a time, the residual leaves in 4-byte pieces that
// complete whole sectors within the warp's store sequence.
struct DctBytes {
  int8_t m[32][32];
};
constexpr DctBytes make_dct_bytes() {
  DctBytes t{};
  for (int j = 0; j < 32; j++)
    for (int i = 0; i < 32; i++) t.m[j][i] = (int8_t)dct32(j, i);
  return t;
}
__device__ const DctBytes kDctBytes = make_dct_bytes();

__device__ __forceinline__ void mma_k32(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, int a_unsigned,
                                        int b_unsigned) {
  if (b_unsigned)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else if (a_unsigned)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k16(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, int a_unsigned, int b_unsigned) {
  if (b_unsigned)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  else if (a_unsigned)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// high / low bytes of four 16-bit values held as two packed pairs
__device__ __forceinline__ uint32_t hi_bytes(uint32_t p01, uint32_t p23) { return __byte_perm(p01, p23, 0x7531); }
__device__ __forceinline__ uint32_t lo_bytes(uint32_t p01, uint32_t p23) { return __byte_perm(p01, p23, 0x6420); }

// two s32 values saturated to s16 and packed (lo in the low half): clip16 x 2 + pack in one instruction
__device__ __forceinline__ uint32_t sat_pack16(int lo, int hi) {
  uint32_t d;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
  return d;
}

// 8.6.4.2 scaling of the eight levels of one 16-byte piece: m = the piece's scaling factors (8 bytes), ms = factor * levelScale.
// WIDE: some level of the block may overflow the 32-bit product (decided once per block from the OR of the magnitudes).
template <int BD_SHIFT, bool WIDE, bool TSKIP>
__device__ __forceinline__ uint4 scale_piece(uint4 raw, uint2 mm, int scale) {
  uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int v2[2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int lvl = hh ? (int)w[k] >> 16 : (int)(int16_t)(w[k] & 0xffffu);
      const int mi = 2 * k + hh;
      const int ms = (int)(((mi < 4 ? mm.x : mm.y) >> (8 * (mi & 3))) & 0xffu) * scale;
      int v;
      if (!WIDE) {
        v = (lvl * ms + (1 << (BD_SHIFT - 1))) >> BD_SHIFT;
      } else {
        long long p = ((long long)lvl * ms + (1ll << (BD_SHIFT - 1))) >> BD_SHIFT;
        v = (int)min(32767ll, max(-32768ll, p));
      }
      if (TSKIP) v = ((clip16(v) << 7) + 2048) >> 12;  // transform skip: the rotation of 8.6.4.2 on the scaled coefficient
      v2[hh] = v;
    }
    w[k] = sat_pack16(v2[0], v2[1]);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// do all products |level| * factor of a block fit 31 bits?  `mag`: OR over the block of the levels' magnitudes-or-complements
__device__ __forceinline__ bool products_fit(uint32_t mag, int max_ms) { return (32 - __clz(mag)) + (32 - __clz((uint32_t)max_ms)) <= 30; }
// OR of |x|-ish bit patterns of the two 16-bit halves of w (x ^ (x >> 15) has the magnitude's leading bit)
__device__ __forceinline__ uint32_t mag_bits(uint32_t w) {
  const int lo = (int)(int16_t)(w & 0xffffu), hi = (int)w >> 16;
  return (uint32_t)(lo ^ (lo >> 31)) | (uint32_t)(hi ^ (hi >> 31));
}

template <int N>
__global__ void __launch_bounds__(kWarpsPerCta * 32, N == 32 ? 4 : 8) transform_mma_kernel(Arenas A, int cls) {
  constexpr int LOG2 = N == 16 ? 4 : 5;
  constexpr int NT = N / 8;    // 8-column tiles
  constexpr int MT = N / 16;   // 16-row tiles
  constexpr int SB = 2 * N + 16;  // row stride of the staging tile in bytes: 8 consecutive rows fall into distinct banks
  constexpr int VEC = N * N / 8 / 32;  // 16-byte pieces of a block per lane
  __shared__ __align__(16) unsigned char tile_all[kWarpsPerCta][N * SB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t count = A.list_count[cls * LIST_COUNT_STRIDE];
  const uint2_t* list = A.tu_list + A.list_off[cls];
  unsigned char* tile = tile_all[warp];
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  // M in the permuted contraction order: cb[h][n] = bytes M[j][8n + g] for j = 16h + {2t, 2t+1, 8+2t, 9+2t}
  uint32_t cb[N / 16][NT];
#pragma unroll
  for (int h = 0; h < N / 16; h++)
#pragma unroll
    for (int n = 0; n < NT; n++) {
      uint32_t v = 0;
#pragma unroll
      for (int s4 = 0; s4 < 4; s4++) {
        const int j = 16 * h + 2 * t + (s4 & 1) + 8 * (s4 >> 1);
        v |= (uint32_t)(uint8_t)kDctBytes.m[j * (32 / N)][8 * n + g] << (8 * s4);
      }
      cb[h][n] = v;
    }
  for (uint32_t idx = blockIdx.x * kWarpsPerCta + warp; idx < count; idx += gridDim.x * kWarpsPerCta) {
    const Item item = fetch_item(A, list[idx]);
    uint4* blk = reinterpret_cast<uint4*>(item.blk);
    uint4 raw[VEC];
#pragma unroll
    for (int q = 0; q < VEC; q++) raw[q] = blk[q * 32 + lane];
    const bool tskip = item.tskip != 0;
    uint32_t other = 0, mag = 0;
#pragma unroll
    for (int q = 0; q < VEC; q++) {
      other |= (q == 0 && lane == 0 ? raw[q].x & 0xffff0000u : raw[q].x) | raw[q].y | raw[q].z | raw[q].w;
      mag |= mag_bits(raw[q].x) | mag_bits(raw[q].y) | mag_bits(raw[q].z) | mag_bits(raw[q].w);
    }
    const int scale = item_scale(item);
    const uint8_t* m = nullptr;
    if (item.pp->scaling_enabled && !tskip) {
      const ScalingSet* sc = A.scaling + item.pp->scaling_set;
      m = N == 16 ? sc->f16[item.cidx] : sc->f32[item.cidx];
    }
    constexpr int BD_SHIFT = LOG2 + 3;  // BitDepth + log2(nTbS) - 5, 8-bit
    if (!__any_sync(0xffffffffu, other != 0) && !tskip) {
      // DC only: both passes of a lone DC term multiply by 64, every sample of the block gets the same value
      const int lvl = (int)(int16_t)(__shfl_sync(0xffffffffu, raw[0].x, 0) & 0xffffu);
      const long long p = ((long long)lvl * ((m ? (int)m[0] : 16) * scale) + (1ll << (BD_SHIFT - 1))) >> BD_SHIFT;
      const int dc = (int)min(32767ll, max(-32768ll, p));
      const int t1 = clip16((64 * dc + 64) >> 7);
      const uint32_t vv = ((uint32_t)clip16((64 * t1 + 2048) >> 12) & 0xffffu) * 0x10001u;
#pragma unroll
      for (int q = 0; q < VEC; q++) blk[q * 32 + lane] = make_uint4(vv, vv, vv, vv);
      continue;
    }
    // ---- scaling (8.6.4.2), in the layout of the load: lane owns 8 consecutive coefficients of a row per piece ----
    const bool narrow = products_fit(__reduce_or_sync(0xffffffffu, mag), (m ? 255 : 16) * scale);
    if (tskip) {  // warp-uniform (one block per warp step): no transform, the scaled coefficients rotated
#pragma unroll
      for (int q = 0; q < VEC; q++) blk[q * 32 + lane] = scale_piece<BD_SHIFT, true, true>(raw[q], make_uint2(0x10101010u, 0x10101010u), scale);
      continue;
    }
#pragma unroll
    for (int q = 0; q < VEC; q++) {
      const int e0 = (q * 32 + lane) * 8;  // element index of the piece
      uint2 mm = make_uint2(0x10101010u, 0x10101010u);
      if (m) mm = *reinterpret_cast<const uint2*>(m + e0);
      const uint4 sc4 = narrow ? scale_piece<BD_SHIFT, false, false>(raw[q], mm, scale) : scale_piece<BD_SHIFT, true, false>(raw[q], mm, scale);
      *reinterpret_cast<uint4*>(tile + (e0 / N) * SB + (e0 % N) * 2) = sc4;
    }
    __syncwarp();
    // ---- column pass: T = M^T X ----
    int acc[MT][NT][4];
#pragma unroll
    for (int n = 0; n < NT; n++) {
      uint32_t r[4];
      if (N == 32) {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(tile_s + (uint32_t)(lane * SB + n * 16)));
      } else {
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                     : "=r"(r[0]), "=r"(r[1])
                     : "r"(tile_s + (uint32_t)((lane & 15) * SB + n * 16)));
        r[2] = r[3] = 0;
      }
      const uint32_t h0 = hi_bytes(r[0], r[1]), l0 = lo_bytes(r[0], r[1]);
      const uint32_t h1 = hi_bytes(r[2], r[3]), l1 = lo_bytes(r[2], r[3]);
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
#pragma unroll
        for (int k = 0; k < 4; k++) acc[mt][n][k] = 0;
        if (N == 32) {
          mma_k32(acc[mt][n], cb[0][2 * mt], cb[0][2 * mt + 1], cb[1][2 * mt], cb[1][2 * mt + 1], h0, h1, 0, 0);
#pragma unroll
          for (int k = 0; k < 4; k++) acc[mt][n][k] = acc[mt][n][k] * 256 + 64;  // the pass's rounding offset rides along
          mma_k32(acc[mt][n], cb[0][2 * mt], cb[0][2 * mt + 1], cb[1][2 * mt], cb[1][2 * mt + 1], l0, l1, 0, 1);
        } else {
          mma_k16(acc[mt][n], cb[0][2 * mt], cb[0][2 * mt + 1], h0, 0, 0);
#pragma unroll
          for (int k = 0; k < 4; k++) acc[mt][n][k] = acc[mt][n][k] * 256 + 64;
          mma_k16(acc[mt][n], cb[0][2 * mt], cb[0][2 * mt + 1], l0, 0, 1);
        }
      }
    }
    // ---- row pass: R = T M, A fragments straight from the accumulators ----
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      uint32_t ah[4], al[4];  // a0: row g, first half of the contraction; a1: row g + 8; a2, a3: second half
#pragma unroll
      for (int hf = 0; hf < N / 16; hf++)
#pragma unroll
        for (int up = 0; up < 2; up++) {
          const uint32_t p01 = sat_pack16(acc[mt][2 * hf][2 * up] >> 7, acc[mt][2 * hf][2 * up + 1] >> 7);
          const uint32_t p23 = sat_pack16(acc[mt][2 * hf + 1][2 * up] >> 7, acc[mt][2 * hf + 1][2 * up + 1] >> 7);
          ah[2 * hf + up] = hi_bytes(p01, p23);
          al[2 * hf + up] = lo_bytes(p01, p23);
        }
#pragma unroll
      for (int n = 0; n < NT; n++) {
        int d[4] = {0, 0, 0, 0};
        if (N == 32) {
          mma_k32(d, ah[0], ah[1], ah[2], ah[3], cb[0][n], cb[1][n], 0, 0);
#pragma unroll
          for (int k = 0; k < 4; k++) d[k] = d[k] * 256 + 2048;
          mma_k32(d, al[0], al[1], al[2], al[3], cb[0][n], cb[1][n], 1, 0);
        } else {
          mma_k16(d, ah[0], ah[1], cb[0][n], 0, 0);
#pragma unroll
          for (int k = 0; k < 4; k++) d[k] = d[k] * 256 + 2048;
          mma_k16(d, al[0], al[1], cb[0][n], 1, 0);
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(item.blk);
        out[((16 * mt + g) * N + 8 * n + 2 * t) >> 1] = sat_pack16(d[0] >> 12, d[1] >> 12);
        out[((16 * mt + g + 8) * N + 8 * n + 2 * t) >> 1] = sat_pack16(d[2] >> 12, d[3] >> 12);
      }
    }
    __syncwarp();
  }
}

// 8x8 blocks, four per warp step (a block is 128 contiguous bytes; a lane loads one 16-byte row of one block).  Two blocks
// share an MMA: the contraction index is (block, j), A is
// block-diagonal [M^T 0; 0 M^T] for the column pass, and in the row pass each block's T occupies its own eight rows of A with
// the other half of the contraction zero.  One constant register per thread.
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) transform_mma8_kernel(Arenas A, int cls) {
  __shared__ __align__(16) unsigned char tile_all[kWarpsPerCta][4 * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t count = A.list_count[cls * LIST_COUNT_STRIDE];
  const uint2_t* list = A.tu_list + A.list_off[cls];
  unsigned char* tile = tile_all[warp];
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  // bytes M8[2t][g], M8[2t+1][g] (M8[j][i] = dct32(4j, i)) in the low / high half of the contraction quad
  const uint32_t c2 = (uint32_t)(uint8_t)kDctBytes.m[4 * (2 * t)][g] | ((uint32_t)(uint8_t)kDctBytes.m[4 * (2 * t + 1)][g] << 8);
  const uint32_t c_lo = c2, c_hi = c2 << 16;
  const uint32_t steps = (count + 3) / 4;
  const int kb = lane >> 3;  // block of the step this lane loads a row of
  for (uint32_t step = blockIdx.x * kWarpsPerCta + warp; step < steps; step += gridDim.x * kWarpsPerCta) {
    const uint32_t idx = step * 4 + kb;
    const bool active = idx < count;
    uint2_t it;
    it.x = it.y = 0;
    if (active) it = list[idx];
    const Item item = fetch_item(A, it);  // item 0 of tile 0 for the idle lanes of the last step: read, never written
    uint4* row = reinterpret_cast<uint4*>(item.blk) + (lane & 7);
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (active) raw = *row;
    const int scale = item_scale(item);
    const bool tskip = active && item.tskip;
    uint2 mm = make_uint2(0x10101010u, 0x10101010u);
    if (item.pp->scaling_enabled && !tskip) mm = *reinterpret_cast<const uint2*>((A.scaling + item.pp->scaling_set)->f8[item.cidx] + (lane & 7) * 8);
    const uint32_t mag = mag_bits(raw.x) | mag_bits(raw.y) | mag_bits(raw.z) | mag_bits(raw.w);
    const bool narrow = products_fit(__reduce_or_sync(0xffffffffu, mag), __reduce_max_sync(0xffffffffu, (item.pp->scaling_enabled ? 255 : 16) * scale));
    const bool any_tskip = __any_sync(0xffffffffu, tskip);
    uint4 sc4;
    if (any_tskip) {  // rare: the skipped blocks are finished here, the others take the wide form
      sc4 = tskip ? scale_piece<6, true, true>(raw, mm, scale) : scale_piece<6, true, false>(raw, mm, scale);
      if (tskip) {
        *row = sc4;
        sc4 = make_uint4(0u, 0u, 0u, 0u);
      }
    } else {
      sc4 = narrow ? scale_piece<6, false, false>(raw, mm, scale) : scale_piece<6, true, false>(raw, mm, scale);
    }
    *reinterpret_cast<uint4*>(tile + lane * 16) = sc4;
    __syncwarp();
    uint32_t r[4];  // block b: X_b[2t][g], X_b[2t+1][g]
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(tile_s + (uint32_t)(lane * 16)));
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {  // blocks 2 pr, 2 pr + 1
      int acc[4] = {0, 0, 0, 0};
      mma_k16(acc, c_lo, c_hi, hi_bytes(r[2 * pr], r[2 * pr + 1]), 0, 0);
#pragma unroll
      for (int k = 0; k < 4; k++) acc[k] = acc[k] * 256 + 64;
      mma_k16(acc, c_lo, c_hi, lo_bytes(r[2 * pr], r[2 * pr + 1]), 0, 1);
      // T_b[g][2t], T_b[g][2t+1] (b = 2 pr: acc[0..1], b = 2 pr + 1: acc[2..3]) -> A of the row pass, second half zero
      const uint32_t p0 = sat_pack16(acc[0] >> 7, acc[1] >> 7), p1 = sat_pack16(acc[2] >> 7, acc[3] >> 7);
      int d[4] = {0, 0, 0, 0};
      mma_k16(d, hi_bytes(p0, 0u), hi_bytes(p1, 0u), c_lo, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; k++) d[k] = d[k] * 256 + 2048;
      mma_k16(d, lo_bytes(p0, 0u), lo_bytes(p1, 0u), c_lo, 1, 0);
      // R_b[g][2t], R_b[g][2t+1]: the lanes of a quad complete a 16-byte row, the warp a whole 128-byte block
#pragma unroll
      for (int b2 = 0; b2 < 2; b2++) {
        const int b = 2 * pr + b2;
        const bool b_active = step * 4 + b < count;
        const bool b_tskip = (__ballot_sync(0xffffffffu, tskip) >> (8 * b)) & 1u;
        uint32_t* out = reinterpret_cast<uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)item.blk, 8 * b));
        if (b_active && !b_tskip) out[(g * 8 + 2 * t) >> 1] = sat_pack16(d[2 * b2] >> 12, d[2 * b2 + 1] >> 12);
      }
    }
    __syncwarp();
  }
}

// ---- n == 4: one lane per block -----------------------------------------------------------------------------
__device__ __forceinline__ void idct4(const int (&x)[4], int (&y)[4]) {
  const int e0 = 64 * (x[0] + x[2]), e1 = 64 * (x[0] - x[2]);
  const int o0 = 83 * x[1] + 36 * x[3], o1 = 36 * x[1] - 83 * x[3];
  y[0] = e0 + o0;
  y[1] = e1 + o1;
  y[2] = e1 - o1;
  y[3] = e0 - o0;
}

// DST: luma 4x4 (intra blocks use the DST-VII, 8.6.4.2); otherwise the 4-point DCT of the chroma blocks.
template <bool DST>
__global__ void __launch_bounds__(256) transform4_kernel(Arenas A, int cls) {
  const uint32_t count = A.list_count[cls * LIST_COUNT_STRIDE];
  const uint2_t* list = A.tu_list + A.list_off[cls];
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
    const Item item = fetch_item(A, list[idx]);
    int16_t* blk = item.blk;
    const uint4 in0 = reinterpret_cast<const uint4*>(blk)[0], in1 = reinterpret_cast<const uint4*>(blk)[1];
    const uint32_t raw[8] = {in0.x, in0.y, in0.z, in0.w, in1.x, in1.y, in1.z, in1.w};
    const int scale = item_scale(item);
    uint32_t mraw[4] = {0x10101010u, 0x10101010u, 0x10101010u, 0x10101010u};
    if (item.pp->scaling_enabled) {
      const uint4 mv = *reinterpret_cast<const uint4*>((A.scaling + item.pp->scaling_set)->f4[item.cidx]);
      mraw[0] = mv.x, mraw[1] = mv.y, mraw[2] = mv.z, mraw[3] = mv.w;
    }
    int d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int lvl = (int)(int16_t)((raw[i >> 1] >> (16 * (i & 1))) & 0xffffu);
      const int ms = (int)((mraw[i >> 2] >> (8 * (i & 3))) & 0xffu) * scale;
      int v;
      if ((unsigned)(lvl + 255) <= 510u) {  // |level| <= 255: 32-bit product (bdShift = 5 for 4x4, 8-bit)
        v = (lvl * ms + 16) >> 5;
      } else {
        long long p = ((long long)lvl * ms + 16) >> 5;
        v = (int)min(32767ll, max(-32768ll, p));
      }
      d[i] = min(32767, max(-32768, v));
    }
    int out[16];
    if (item.tskip) {
#pragma unroll
      for (int i = 0; i < 16; i++) out[i] = ((d[i] << 7) + 2048) >> 12;
    } else {
      int t[16];
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {  // columns
        const int x[4] = {d[cc], d[4 + cc], d[8 + cc], d[12 + cc]};
        int y[4];
        if (DST) dst4(x, y);
        else idct4(x, y);
#pragma unroll
        for (int i = 0; i < 4; i++) t[i * 4 + cc] = clip16((y[i] + 64) >> 7);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {  // rows
        const int x[4] = {t[r * 4], t[r * 4 + 1], t[r * 4 + 2], t[r * 4 + 3]};
        int y[4];
        if (DST) dst4(x, y);
        else idct4(x, y);
#pragma unroll
        for (int i = 0; i < 4; i++) out[r * 4 + i] = clip16((y[i] + 2048) >> 12);
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; i++) pk[i] = ((uint32_t)out[2 * i] & 0xffffu) | ((uint32_t)out[2 * i + 1] << 16);
    reinterpret_cast<uint4*>(blk)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    reinterpret_cast<uint4*>(blk)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

}  // namespace

// Worst-case list sizes for `total_tu` tu_map entries: class offsets (elements) and the total, for the batch's allocation.
size_t transform_list_layout(size_t total_tu, uint32_t off[LIST_CLASSES]) {
  const size_t cap[LIST_CLASSES] = {total_tu / 64 + 1, total_tu / 16 + total_tu / 32 + 1, total_tu / 4 + total_tu / 8 + 1,
                                    total_tu + 1, total_tu / 2 + 1};
  size_t o = 0;
  for (int k = 0; k < LIST_CLASSES; k++) {
    off[k] = (uint32_t)o;
    o += cap[k];
  }
  return o;
}

cudaError_t launch_transform(const Arenas& A, uint32_t max_tu_per_tile, int max_log2_tb, int n_sm, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(A.list_count, 0, LIST_CLASSES * LIST_COUNT_STRIDE * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  const uint32_t bpt = (max_tu_per_tile + kListPerCta - 1) / kListPerCta;
  tu_list_kernel<<<A.n_tiles * bpt, kListThreads, 0, stream>>>(A, bpt);
  // the list lengths stay on the device: fixed grids sized to the machine stride over them
  const size_t total = (size_t)A.n_tiles * max_tu_per_tile;
  auto grid = [&](size_t max_steps, int resident) {
    return (unsigned)std::max<size_t>(1, std::min<size_t>((max_steps + kWarpsPerCta - 1) / kWarpsPerCta, (size_t)n_sm * resident));
  };
  // HEIC_B200_TRANSFORM_MMA=0: the register butterflies for 16x16 / 32x32 as well (A/B against the tensor-core form)
  static const bool use_mma = []() { const char* e = getenv("HEIC_B200_TRANSFORM_MMA"); return !e || atoi(e) != 0; }();
  if (use_mma) {
    if (max_log2_tb >= 5) transform_mma_kernel<32><<<grid(total / 64 + 1, 4), kWarpsPerCta * 32, 0, stream>>>(A, LIST_32);
    if (max_log2_tb >= 4) transform_mma_kernel<16><<<grid(total / 16 + 1, 8), kWarpsPerCta * 32, 0, stream>>>(A, LIST_16);
  } else {
    if (max_log2_tb >= 5) transform_kernel<32><<<grid(total / 64 + 1, 3), kWarpsPerCta * 32, 0, stream>>>(A, LIST_32);
    if (max_log2_tb >= 4) transform_kernel<16><<<grid(total / 32 + 1, 4), kWarpsPerCta * 32, 0, stream>>>(A, LIST_16);
  }
  if (use_mma) transform_mma8_kernel<<<grid(total / 16 + 1, 8), kWarpsPerCta * 32, 0, stream>>>(A, LIST_8);
  else transform_kernel<8><<<grid(total / 16 + 1, 6), kWarpsPerCta * 32, 0, stream>>>(A, LIST_8);
  transform4_kernel<true><<<grid(total / 32 + 1, 8), 256, 0, stream>>>(A, LIST_4Y);
  transform4_kernel<false><<<grid(total / 64 + 1, 8), 256, 0, stream>>>(A, LIST_4C);
  return cudaGetLastError();
}
int transform_launches(int max_log2_tb) { return 4 + (max_log2_tb >= 4 ? 1 : 0) + (max_log2_tb >= 5 ? 1 : 0); }

}  // namespace dev
}  // namespace heic
