// Stage 2: scaling (H.265 8.6.4.2 with the 7.4.5 scaling factors) + inverse DST-4 / DCT-4..32 (8.6.4.2),
// in place on the coefficient arena.  The reference has none of this (slice.rs:253-255 is todo!()).
//
// Small integer butterflies, HBM-bound.  One launch per (component kind, transform size) so that every CTA of
// a launch runs the same straight-line butterfly and the instruction cache holds it (a single kernel with all
// sizes was instruction-fetch bound).  A warp owns 64 consecutive tu_map entries (a 32x32 luma area and its
// two 16x16 chroma areas) and sweeps the aligned slots of the launch's size:
//   n >= 8: 32/n transform blocks at a time; a lane runs one n-point 1-D transform per pass (columns, then
//           rows) as an even/odd partial butterfly in registers; the passes are joined through a padded
//           (conflict-free) shared-memory tile and the residual goes back to HBM in lane-contiguous words;
//   n == 4: one lane per block, both passes in registers, 32 blocks per warp step, 32-byte loads/stores.
// Only coded blocks (cbf = 1) are read or written, so DRAM traffic is 2 B in + 2 B out per coded sample.
#include <cuda_runtime.h>

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

struct WarpCtx {
  const PicParams* pp;
  const TileParams* tp;
  const ScalingSet* sc;
  int16_t* tmp;   // this warp's padded transpose tile
  const uint32_t* tu;  // the region's 64 tu_map words
  int n_entries;       // valid entries of the region (< 64 only at the end of a CTB-16 picture)
  int lane;
};

// One transform size of one colour component over the warp's region.
//   luma   (CIDX 0): slots of (N/4)^2 tu_map entries; block at coeff + slot * N*N
//   chroma (CIDX>0): chroma NxN belongs to the luma TU of size 2N (or to blkIdx 3 of a split 8x8 when N = 4)
template <int N, int CIDX>
__device__ __forceinline__ void run_size(const WarpCtx& c, int16_t* coeff /* region base of this component */) {
  constexpr int K = 32 / N;                      // blocks per batch
  constexpr int LOG2 = N == 4 ? 2 : N == 8 ? 3 : N == 16 ? 4 : 5;
  constexpr int EPB = CIDX == 0 ? (N / 4) * (N / 4) : (N / 2) * (N / 2);  // tu_map entries per block slot
  constexpr int SLOTS = 64 / EPB;
  constexpr int S = N + 2;                       // padded row stride (int16) of the transpose tile
  constexpr bool DST = (CIDX == 0 && N == 4);
  const int k = c.lane / N, col = c.lane % N;
  for (int batch = 0; batch < (SLOTS + K - 1) / K; batch++) {
    const int slot = batch * K + k;
    bool active = false;
    uint32_t w;
    {
      const bool has_slot = slot < SLOTS && slot * EPB < c.n_entries;
      const int entry = has_slot ? slot * EPB : 0;
      w = c.tu[entry];  // lanes of one block read the same word (broadcast)
      const uint32_t cbf_bit = CIDX == 0 ? TU_CBF_Y : (CIDX == 1 ? TU_CBF_CB : TU_CBF_CR);
      if (CIDX == 0) active = (w & TU_ORIGIN) && tu_log2(w) == LOG2 && (w & cbf_bit);
      else active = (w & TU_ORIGIN) && tu_log2(w) == LOG2 + 1 && (w & cbf_bit);
      active = active && has_slot;
    }
    if (!__any_sync(0xffffffffu, active)) continue;

    int16_t* blk = coeff + (size_t)slot * (N * N);
    int16_t* t = c.tmp + k * (N * S);
    const bool tskip = active && tu_tskip(w, CIDX);
    // ---- scaling + first (column) pass -------------------------------------------------------
    int x[N], y[N];
    int nz_rows = 0;
    // raw levels first: only the rows up to the last non-zero one (warp-uniform bound) are scaled afterwards
#pragma unroll
    for (int j = 0; j < N; j++) {
      x[j] = active ? (int)blk[j * N + col] : 0;
      nz_rows = x[j] ? j + 1 : nz_rows;
    }
    // extents of the non-zero coefficients, uniform over the warp so the butterfly variant is too
    int nz1 = nz_rows;
#pragma unroll
    for (int o = 16; o; o >>= 1) nz1 = max(nz1, __shfl_xor_sync(0xffffffffu, nz1, o));
    const uint32_t col_mask = __ballot_sync(0xffffffffu, nz_rows > 0);
    // highest non-zero column index + 1 within any block of the batch
    int nz2 = 0;
#pragma unroll
    for (int g = 0; g < K; g++) {
      uint32_t mg = (col_mask >> (g * N)) & (N == 32 ? 0xffffffffu : ((1u << N) - 1u));
      nz2 = max(nz2, 32 - __clz(mg));
    }
    if (active) {
      int qp = (int)tu_qp(w);
      if (CIDX) {
        int qpi = qp + (CIDX == 1 ? c.pp->pps_cb_qp_offset + c.tp->slice_cb_qp_offset
                                  : c.pp->pps_cr_qp_offset + c.tp->slice_cr_qp_offset);
        qpi = min(57, max(0, qpi));
        qp = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : kChromaQp[qpi - 30]);
      }
      const int scale = (int)kLevelScale[qp % 6] << (qp / 6);
      const uint8_t* m = nullptr;
      if (c.pp->scaling_enabled && !(tskip && N > 4))
        m = N == 4 ? c.sc->f4[CIDX] : N == 8 ? c.sc->f8[CIDX] : N == 16 ? c.sc->f16[CIDX] : c.sc->f32[CIDX];
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
    if (tskip) {
#pragma unroll
      for (int i = 0; i < N; i++) y[i] = x[i];  // passed through; the rotation happens in the second pass
    } else {
      transform_1d<N, DST>(x, y, nz1);
#pragma unroll
      for (int i = 0; i < N; i++) y[i] = clip16((y[i] + 64) >> 7);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N; i++) t[i * S + col] = (int16_t)y[i];
    __syncwarp();
    // ---- second (row) pass: lane `col` now owns row `col` of its block -----------------------------
    {
      const int16_t* row = t + col * S;
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        uint32_t two = *reinterpret_cast<const uint32_t*>(row + j);
        x[j] = (int)(int16_t)(two & 0xffffu);
        x[j + 1] = (int)(int16_t)(two >> 16);
      }
    }
    if (tskip) {
#pragma unroll
      for (int i = 0; i < N; i++) y[i] = ((x[i] << 7) + 2048) >> 12;
    } else {
      transform_1d<N, DST>(x, y, nz2);
#pragma unroll
      for (int i = 0; i < N; i++) y[i] = clip16((y[i] + 2048) >> 12);
    }
    __syncwarp();
    {
      int16_t* row = t + col * S;
#pragma unroll
      for (int j = 0; j < N; j += 2)
        *reinterpret_cast<uint32_t*>(row + j) = ((uint32_t)y[j] & 0xffffu) | ((uint32_t)y[j + 1] << 16);
    }
    __syncwarp();
    // ---- coalesced store of the residual blocks (pairs of int16) ----------------------------------
    {
      constexpr int PAIRS = N * N / 2;            // per block
      constexpr int TOTAL = K * PAIRS;            // per batch, lane-consecutive
#pragma unroll 4
      for (int p = c.lane; p < TOTAL; p += 32) {
        const int g = p / PAIRS, q = p % PAIRS;
        const int r = q / (N / 2), cp = q % (N / 2);
        const bool g_active = __shfl_sync(0xffffffffu, active ? 1 : 0, g * N) != 0;
        const int g_slot = batch * K + g;
        if (g_active) {
          uint32_t v = *reinterpret_cast<const uint32_t*>(c.tmp + g * (N * S) + r * S + 2 * cp);
          *reinterpret_cast<uint32_t*>(coeff + (size_t)g_slot * (N * N) + 2 * q) = v;
        }
      }
    }
    __syncwarp();
  }
}

constexpr int kWarpsPerCta = 8;
constexpr uint32_t kRegionsPerCta = 64;
constexpr int kTmpPerWarp = 32 * 34;  // int16 elements: one 32x32 block with padded rows

// ---- n >= 8: one launch per (CIDX kind, N) ------------------------------------------------------------------
// KIND 0: luma, KIND 1: chroma (Cb then Cr)
template <int N, int KIND>
__global__ void __launch_bounds__(kWarpsPerCta * 32) transform_kernel(Arenas A) {
  __shared__ __align__(16) int16_t tmp_all[kWarpsPerCta][kTmpPerWarp];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0) return;
  if (KIND == 1 && !pp->chroma) return;
  WarpCtx c;
  c.pp = pp;
  c.tp = tp;
  c.sc = A.scaling + pp->scaling_set;
  c.tmp = tmp_all[warp];
  c.lane = lane;
  // a CTA sweeps kRegionsPerCta consecutive regions, one per warp at a time (fewer, longer-lived CTAs)
  const uint32_t r_end = min((blockIdx.y + 1) * kRegionsPerCta, ((uint32_t)pp->n_tu + 63u) / 64u);
  for (uint32_t region = blockIdx.y * kRegionsPerCta + warp; region < r_end; region += kWarpsPerCta) {
    c.tu = A.tu_map + tp->tu_off + (size_t)region * 64;
    c.n_entries = min(64, pp->n_tu - (int)region * 64);
    if (KIND == 0) {
      run_size<N, 0>(c, A.coeff + tp->coeff_off[0] + (size_t)region * 1024);
    } else {
      run_size<N, 1>(c, A.coeff + tp->coeff_off[1] + (size_t)region * 256);
      run_size<N, 2>(c, A.coeff + tp->coeff_off[2] + (size_t)region * 256);
    }
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

template <int CIDX>
__device__ __forceinline__ void block4(int16_t* blk, uint32_t w, const PicParams* pp, const TileParams* tp, const ScalingSet* sc) {
  const uint4 in0 = reinterpret_cast<const uint4*>(blk)[0], in1 = reinterpret_cast<const uint4*>(blk)[1];
  const uint32_t raw[8] = {in0.x, in0.y, in0.z, in0.w, in1.x, in1.y, in1.z, in1.w};
  int qp = (int)tu_qp(w);
  if (CIDX) {
    int qpi = qp + (CIDX == 1 ? pp->pps_cb_qp_offset + tp->slice_cb_qp_offset : pp->pps_cr_qp_offset + tp->slice_cr_qp_offset);
    qpi = min(57, max(0, qpi));
    qp = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : kChromaQp[qpi - 30]);
  }
  const int scale = (int)kLevelScale[qp % 6] << (qp / 6);
  uint32_t mraw[4] = {0x10101010u, 0x10101010u, 0x10101010u, 0x10101010u};
  if (pp->scaling_enabled) {
    const uint4 mv = *reinterpret_cast<const uint4*>(sc->f4[CIDX]);
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
  if (tu_tskip(w, CIDX)) {
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = ((d[i] << 7) + 2048) >> 12;
  } else {
    int t[16];
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {  // columns
      const int x[4] = {d[cc], d[4 + cc], d[8 + cc], d[12 + cc]};
      int y[4];
      if (CIDX == 0) dst4(x, y);
      else idct4(x, y);
#pragma unroll
      for (int i = 0; i < 4; i++) t[i * 4 + cc] = clip16((y[i] + 64) >> 7);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {  // rows
      const int x[4] = {t[r * 4], t[r * 4 + 1], t[r * 4 + 2], t[r * 4 + 3]};
      int y[4];
      if (CIDX == 0) dst4(x, y);
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

// KIND 0: luma 4x4 TUs (a thread per tu_map entry); KIND 1: chroma 4x4 (a thread per 8x8 luma area and component)
template <int KIND>
__global__ void __launch_bounds__(256) transform4_kernel(Arenas A) {
  const uint32_t tile = blockIdx.x;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0) return;
  const ScalingSet* sc = A.scaling + pp->scaling_set;
  const uint32_t* tu = A.tu_map + tp->tu_off;
  const uint32_t i = blockIdx.y * blockDim.x + threadIdx.x;
  if (KIND == 0) {
    if (i >= (uint32_t)pp->n_tu) return;
    const uint32_t w = tu[i];
    if ((w & TU_ORIGIN) && tu_log2(w) == 2 && (w & TU_CBF_Y)) block4<0>(A.coeff + tp->coeff_off[0] + (size_t)i * 16, w, pp, tp, sc);
  } else {
    if (!pp->chroma) return;
    const uint32_t slot = i >> 1, comp = i & 1;  // Cb / Cr of one slot on adjacent lanes
    if (slot * 4 >= (uint32_t)pp->n_tu) return;
    uint32_t w = tu[slot * 4];
    bool ok = false;
    if ((w & TU_ORIGIN) && tu_log2(w) == 3) ok = true;             // 8x8 luma TU -> 4x4 chroma
    else if ((w & TU_ORIGIN) && tu_log2(w) == 2) {                  // split 8x8: chroma rides on blkIdx 3
      w = tu[slot * 4 + 3];
      ok = (w & TU_ORIGIN) && (w & TU_HAS_CHROMA);
    }
    if (!ok) return;
    if (comp == 0) {
      if (w & TU_CBF_CB) block4<1>(A.coeff + tp->coeff_off[1] + (size_t)slot * 16, w, pp, tp, sc);
    } else {
      if (w & TU_CBF_CR) block4<2>(A.coeff + tp->coeff_off[2] + (size_t)slot * 16, w, pp, tp, sc);
    }
  }
}

}  // namespace

cudaError_t launch_transform(const Arenas& A, uint32_t max_tu_per_tile, int max_log2_tb, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const uint32_t regions = (max_tu_per_tile + 63u) / 64u;
  const dim3 grid(A.n_tiles, (regions + kRegionsPerCta - 1) / kRegionsPerCta), block(kWarpsPerCta * 32);
  if (max_log2_tb >= 5) {
    transform_kernel<32, 0><<<grid, block, 0, stream>>>(A);
    transform_kernel<16, 1><<<grid, block, 0, stream>>>(A);
  }
  if (max_log2_tb >= 4) {
    transform_kernel<16, 0><<<grid, block, 0, stream>>>(A);
    transform_kernel<8, 1><<<grid, block, 0, stream>>>(A);
  }
  transform_kernel<8, 0><<<grid, block, 0, stream>>>(A);
  transform4_kernel<0><<<dim3(A.n_tiles, (max_tu_per_tile + 255) / 256), 256, 0, stream>>>(A);
  transform4_kernel<1><<<dim3(A.n_tiles, (max_tu_per_tile / 2 + 255) / 256), 256, 0, stream>>>(A);
  return cudaGetLastError();
}
int transform_launches(int max_log2_tb) { return 3 + (max_log2_tb >= 4 ? 2 : 0) + (max_log2_tb >= 5 ? 2 : 0); }

}  // namespace dev
}  // namespace heic
