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

__global__ void __launch_bounds__(kListThreads) tu_list_kernel(Arenas A, uint32_t blocks_per_tile) {
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
  uint32_t cls[kListIter][4];
  unsigned long long mine = 0;  // items of this thread per class, 12 bits each
#pragma unroll
  for (int it = 0; it < kListIter; it++) {
    const uint32_t i0 = first + (uint32_t)it * (4 * kListThreads) + threadIdx.x * 4;
    uint4 wv = make_uint4(0u, 0u, 0u, 0u);
    if (i0 < (uint32_t)pp->n_tu) wv = *reinterpret_cast<const uint4*>(A.tu_map + tp->tu_off + i0);
    const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
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
          uint2_t v;
          v.x = tile;
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
  uint32_t w;
  int cidx;
};
__device__ __forceinline__ Item fetch_item(const Arenas& A, uint2_t it) {
  Item r;
  r.tp = A.tiles + it.x;
  r.pp = A.pics + r.tp->pic;
  const uint32_t entry = it.y & 0x3fffffffu;
  r.cidx = (int)(it.y >> 30);
  r.w = A.tu_map[r.tp->tu_off + entry];
  // luma: 16 coefficients per tu_map entry; chroma: 4, at the entry of the 8x8 luma area the block belongs to
  r.blk = A.coeff + r.tp->coeff_off[r.cidx] + (r.cidx ? (size_t)(entry & ~3u) * 4 : (size_t)entry * 16);
  return r;
}
__device__ __forceinline__ int item_scale(const Item& t) {  // levelScale[qP % 6] << (qP / 6), 8.6.2 / 8.6.4.2
  int qp = (int)tu_qp(t.w);
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
    const bool tskip = active && tu_tskip(item.w, item.cidx);
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
    if (tu_tskip(item.w, item.cidx)) {
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
  if (max_log2_tb >= 5) transform_kernel<32><<<grid(total / 64 + 1, 3), kWarpsPerCta * 32, 0, stream>>>(A, LIST_32);
  if (max_log2_tb >= 4) transform_kernel<16><<<grid(total / 32 + 1, 4), kWarpsPerCta * 32, 0, stream>>>(A, LIST_16);
  transform_kernel<8><<<grid(total / 16 + 1, 6), kWarpsPerCta * 32, 0, stream>>>(A, LIST_8);
  transform4_kernel<true><<<grid(total / 32 + 1, 8), 256, 0, stream>>>(A, LIST_4Y);
  transform4_kernel<false><<<grid(total / 64 + 1, 8), 256, 0, stream>>>(A, LIST_4C);
  return cudaGetLastError();
}
int transform_launches(int max_log2_tb) { return 4 + (max_log2_tb >= 4 ? 1 : 0) + (max_log2_tb >= 5 ? 1 : 0); }

}  // namespace dev
}  // namespace heic
