// Stage 2: scaling (H.265 8.6.4.2 with the 7.4.5 scaling factors) + inverse DST-4 / DCT-4..32 (8.6.4.2),
// in place on the coefficient arena.  The reference has none of this (slice.rs:253-255 is todo!()).
//
// Small integer butterflies, HBM-bound: one warp owns 64 consecutive tu_map entries (a 32x32 luma area,
// 2 KB of coefficients, plus its two 16x16 chroma areas).  For every transform size the warp sweeps the
// aligned slots of that size, 32/n transform blocks at a time: a lane runs one n-point 1-D transform per
// pass (columns, then rows) as an even/odd partial butterfly entirely in registers; the two passes are
// joined through a padded (conflict-free) shared-memory tile.  Only coded blocks (cbf = 1) are read or
// written, so DRAM traffic is 2 B in + 2 B out per coded sample.
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
  uint32_t w0, w1;  // tu_map words of entries lane and lane + 32 of the region
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
      // shuffles are warp-collective: lanes without a slot of this size fetch slot 0 and discard it
      const bool has_slot = slot < SLOTS;
      const int entry = has_slot ? slot * EPB : 0;
      const uint32_t a = __shfl_sync(0xffffffffu, c.w0, entry & 31), b = __shfl_sync(0xffffffffu, c.w1, entry & 31);
      const uint32_t a3 = __shfl_sync(0xffffffffu, c.w0, (entry + 3) & 31), b3 = __shfl_sync(0xffffffffu, c.w1, (entry + 3) & 31);
      w = entry < 32 ? a : b;
      const uint32_t w3 = entry < 32 ? a3 : b3;
      const uint32_t cbf_bit = CIDX == 0 ? TU_CBF_Y : (CIDX == 1 ? TU_CBF_CB : TU_CBF_CR);
      if (CIDX == 0) {
        active = (w & TU_ORIGIN) && tu_log2(w) == LOG2 && (w & cbf_bit);
      } else if (N > 4) {
        active = (w & TU_ORIGIN) && tu_log2(w) == LOG2 + 1 && (w & cbf_bit);
      } else {
        // 4x4 chroma: luma TU 8x8 at this entry, or the fourth 4x4 luma TU of a split 8x8 (entry + 3)
        if ((w & TU_ORIGIN) && tu_log2(w) == 3) {
          active = (w & cbf_bit) != 0;
        } else if ((w & TU_ORIGIN) && tu_log2(w) == 2 && (w3 & TU_ORIGIN) && (w3 & TU_HAS_CHROMA)) {
          w = w3;
          active = (w & cbf_bit) != 0;
        }
      }
      active = active && has_slot;
    }
    if (!__any_sync(0xffffffffu, active)) continue;

    int16_t* blk = coeff + (size_t)slot * (N * N);
    int16_t* t = c.tmp + k * (N * S);
    const bool tskip = active && tu_tskip(w, CIDX);
    // ---- scaling + first (column) pass -------------------------------------------------------
    int x[N], y[N];
    int nz_rows = 0;
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
        int lvl = blk[j * N + col];
        int v = 0;
        if (lvl) {
          int mm = m ? (int)m[j * N + col] : 16;
          long long p = (long long)lvl * (mm * scale) + (1ll << (BD_SHIFT - 1));
          p >>= BD_SHIFT;
          v = (int)(p < -32768 ? -32768 : (p > 32767 ? 32767 : p));
          nz_rows = j + 1;
        }
        x[j] = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; j++) x[j] = 0;
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
constexpr int kTmpPerWarp = 32 * 34;  // int16 elements: one 32x32 block with padded rows

__global__ void __launch_bounds__(kWarpsPerCta * 32) transform_kernel(Arenas A) {
  __shared__ __align__(16) int16_t tmp_all[kWarpsPerCta][kTmpPerWarp];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0) return;
  const uint32_t region = blockIdx.y * kWarpsPerCta + warp;
  const uint32_t n_regions = ((uint32_t)pp->n_tu + 63u) / 64u;
  if (region >= n_regions) return;
  const uint32_t* tu = A.tu_map + tp->tu_off + (size_t)region * 64;
  WarpCtx c;
  c.pp = pp;
  c.tp = tp;
  c.sc = A.scaling + pp->scaling_set;
  c.tmp = tmp_all[warp];
  c.lane = lane;
  const uint32_t n_tu = (uint32_t)pp->n_tu;
  c.w0 = region * 64 + lane < n_tu ? tu[lane] : 0u;
  c.w1 = region * 64 + 32 + lane < n_tu ? tu[32 + lane] : 0u;
  const uint32_t any_y = __ballot_sync(0xffffffffu, ((c.w0 | c.w1) & TU_CBF_Y) != 0);
  const uint32_t any_c = __ballot_sync(0xffffffffu, ((c.w0 | c.w1) & (TU_CBF_CB | TU_CBF_CR)) != 0);
  if (any_y) {
    int16_t* y = A.coeff + tp->coeff_off[0] + (size_t)region * 1024;
    uint32_t sizes = 0;  // bit l set: some coded luma block of log2 size l + 2 exists in the region
    {
      uint32_t s = 0;
      if (c.w0 & TU_CBF_Y) s |= 1u << ((c.w0 >> 1) & 3);
      if (c.w1 & TU_CBF_Y) s |= 1u << ((c.w1 >> 1) & 3);
#pragma unroll
      for (int o = 16; o; o >>= 1) s |= __shfl_xor_sync(0xffffffffu, s, o);
      sizes = s;
    }
    if (sizes & 8u) run_size<32, 0>(c, y);
    if (sizes & 4u) run_size<16, 0>(c, y);
    if (sizes & 2u) run_size<8, 0>(c, y);
    if (sizes & 1u) run_size<4, 0>(c, y);
  }
  if (any_c && pp->chroma) {
    int16_t* cb = A.coeff + tp->coeff_off[1] + (size_t)region * 256;
    int16_t* cr = A.coeff + tp->coeff_off[2] + (size_t)region * 256;
    uint32_t sizes = 0;  // by chroma log2 size - 2
    {
      uint32_t s = 0;
      if (c.w0 & (TU_CBF_CB | TU_CBF_CR)) s |= 1u << max(0, (int)((c.w0 >> 1) & 3) - 1);
      if (c.w1 & (TU_CBF_CB | TU_CBF_CR)) s |= 1u << max(0, (int)((c.w1 >> 1) & 3) - 1);
#pragma unroll
      for (int o = 16; o; o >>= 1) s |= __shfl_xor_sync(0xffffffffu, s, o);
      sizes = s;
    }
    if (sizes & 4u) {
      run_size<16, 1>(c, cb);
      run_size<16, 2>(c, cr);
    }
    if (sizes & 2u) {
      run_size<8, 1>(c, cb);
      run_size<8, 2>(c, cr);
    }
    if (sizes & 1u) {
      run_size<4, 1>(c, cb);
      run_size<4, 2>(c, cr);
    }
  }
}

}  // namespace

cudaError_t launch_transform(const Arenas& A, uint32_t max_tu_per_tile, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const uint32_t regions = (max_tu_per_tile + 63u) / 64u;
  dim3 grid(A.n_tiles, (regions + kWarpsPerCta - 1) / kWarpsPerCta);
  transform_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(A);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
