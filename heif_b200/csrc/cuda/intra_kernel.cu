// Stage 3: intra prediction + reconstruction (H.265 8.4.4.2.1-6), absent from the reference
// (slice.rs:253-255 is todo!()).
//
// One CTA per picture (HEIF grid tile); warp r owns CTB rows r, r + R, ... and the rows advance as a
// wavefront (a CTU needs its above-right neighbour, so row r trails row r - 1 by two CTUs; progress
// counters live in shared memory).  Inside a CTU the transform units are inherently sequential (each
// predicts from its neighbours' reconstruction); the 32 lanes split the reference-sample gathering,
// substitution, smoothing and the pixels of the block.  The CTU under construction sits in shared memory
// with a one-sample apron (row above incl. above-right CTU, column to the left) and is written to HBM
// once, coalesced, when complete.
#include <cuda_runtime.h>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

constexpr int kMaxRows = 512;

__device__ const int8_t kIntraPredAngle[35] = {0,   0,   32,  26,  21,  17, 13, 9,  5,  2,  0,  -2, -5, -9, -13, -17, -21, -26,
                                               -32, -26, -21, -17, -13, -9, -5, -2, 0,  2,  5,  9,  13, 17, 21,  26,  32};
__device__ const int16_t kInvAngle[15] = {-4096, -1638, -910, -630, -482, -390, -315, -256,
                                          -315,  -390,  -482, -630, -910, -1638, -4096};  // modes 11..25

__device__ __forceinline__ uint32_t compact1(uint32_t v) {
  v &= 0x55u;
  v = (v | (v >> 1)) & 0x33u;
  v = (v | (v >> 2)) & 0x0fu;
  return v;
}
__device__ __forceinline__ int clip8(int v) { return min(255, max(0, v)); }

struct Ctu {
  // geometry of the picture / CTU
  int w, h, ctb, ctb4, x_ctb, y_ctb;
  // per-warp shared memory
  uint8_t* buf[3];   // sample (x, y) relative to the CTU origin at buf[(y + 1) * stride + x + 16], x, y >= -1
  int stride[3];
  uint8_t* ref;      // reference samples after substitution, s = 0 .. 4n (bottom-left -> corner -> top-right)
  uint8_t* reff;     // after smoothing
  int16_t* refa;     // angular reference array, refa[-32 .. 64]
  uint32_t* dmask;   // reconstructed 4x4 luma blocks of this CTU, bit uy * ctb4 + ux
  int lane;
  int strong_flag;
};

// 6.4.1 z-scan availability restated for a single-slice all-intra picture decoded in wavefront order.
// (px, py): sample position in plane coordinates relative to the CTU origin; sub = 1 for 4:2:0 chroma.
__device__ __forceinline__ bool sample_available(const Ctu& c, int px, int py, int sub) {
  const int lx = px * (1 << sub), ly = py * (1 << sub);
  const int X = c.x_ctb + lx, Y = c.y_ctb + ly;
  if (X < 0 || Y < 0 || X >= c.w || Y >= c.h) return false;
  if (ly < 0) return true;           // CTB row above (incl. above-right CTU: the wavefront guarantees it)
  if (lx < 0) return ly < c.ctb;     // CTU to the left; below this CTB row nothing is decoded yet
  if (lx >= c.ctb || ly >= c.ctb) return false;
  const int bit = (ly >> 2) * c.ctb4 + (lx >> 2);
  return (c.dmask[bit >> 5] >> (bit & 31)) & 1u;
}

// Predicts and reconstructs one n x n block of plane `cidx` at (bx, by) (plane samples, CTU-relative).
__device__ void predict_block(const Ctu& c, int cidx, int bx, int by, int log2, int mode, bool cbf,
                              const int16_t* __restrict__ resid) {
  const int n = 1 << log2, sub = cidx ? 1 : 0, lane = c.lane;
  uint8_t* buf = c.buf[cidx];
  const int stride = c.stride[cidx];
  const int total = 4 * n + 1;
  // ---- 8.4.4.2.2: gather + availability -----------------------------------------------------------
  uint32_t am[5];
#pragma unroll
  for (int m = 0; m < 5; m++) {
    const int s = m * 32 + lane;
    bool a = false;
    if (m * 32 < total) {
      if (s < total) {
        const int dx = s <= 2 * n ? -1 : s - 2 * n - 1;
        const int dy = s <= 2 * n ? 2 * n - 1 - s : -1;
        a = sample_available(c, bx + dx, by + dy, sub);
        c.ref[s] = a ? buf[(by + dy + 1) * stride + bx + dx + 16] : (uint8_t)128;
      }
      am[m] = __ballot_sync(0xffffffffu, a);
    } else {
      am[m] = 0;
    }
  }
  __syncwarp();
  const bool any_avail = (am[0] | am[1] | am[2] | am[3] | am[4]) != 0;
  const bool all_avail = am[0] == 0xffffffffu && (total <= 32 || ((am[1] == 0xffffffffu || total <= 33) && true));
  (void)all_avail;
  if (any_avail) {
    // every unavailable sample takes the nearest available one before it in scan order; if there is none,
    // the first available one after it
#pragma unroll
    for (int m = 0; m < 5; m++) {
      const int s = m * 32 + lane;
      if (m * 32 < total && s < total && !((am[m] >> lane) & 1u)) {
        int src = -1;
        uint32_t bits = am[m] & ((lane == 31) ? 0xffffffffu : ((1u << (lane + 1)) - 1u));
        if (bits) src = m * 32 + 31 - __clz(bits);
#pragma unroll
        for (int k = 4; k >= 0; k--)
          if (k < m && src < 0 && am[k]) src = k * 32 + 31 - __clz(am[k]);
        if (src < 0) {
#pragma unroll
          for (int k = 0; k < 5; k++)
            if (src < 0 && am[k]) src = k * 32 + __ffs(am[k]) - 1;
        }
        c.ref[s] = c.ref[src];
      }
    }
    __syncwarp();
  }
  // ---- 8.4.4.2.3: smoothing (luma only in 4:2:0) ---------------------------------------------------
  const uint8_t* R = c.ref;
  if (cidx == 0 && mode != 1 && n != 4) {
    const int d1 = abs(mode - 26), d2 = abs(mode - 10);
    const int min_dist = min(d1, d2);
    const int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (min_dist > thr) {
      const bool strong = c.strong_flag && n == 32 && abs((int)c.ref[64] + c.ref[128] - 2 * c.ref[96]) < 8 &&
                          abs((int)c.ref[64] + c.ref[0] - 2 * c.ref[32]) < 8;
      for (int s = lane; s < total; s += 32) {
        int v;
        if (s == 0 || s == total - 1) v = c.ref[s];
        else if (strong) {
          if (s == 64) v = c.ref[64];
          else if (s < 64) v = (s * c.ref[64] + (64 - s) * c.ref[0] + 32) >> 6;
          else v = ((128 - s) * c.ref[64] + (s - 64) * c.ref[128] + 32) >> 6;
        } else {
          v = (c.ref[s - 1] + 2 * c.ref[s] + c.ref[s + 1] + 2) >> 2;
        }
        c.reff[s] = (uint8_t)v;
      }
      __syncwarp();
      R = c.reff;
    }
  }
  // left(i) = p[-1][i - 1], top(i) = p[i - 1][-1]; index 0 is the corner
#define LEFT(i) ((int)R[2 * n - (i)])
#define TOP(i) ((int)R[2 * n + (i)])
  int dc = 0, angle = 0;
  bool vertical = false;
  if (mode == 1) {
    int s = 0;
    for (int i = lane; i < n; i += 32) s += LEFT(1 + i) + TOP(1 + i);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    dc = (s + n) >> (log2 + 1);
  } else if (mode >= 2) {
    angle = kIntraPredAngle[mode];
    vertical = mode >= 18;
    int16_t* ra = c.refa;  // ra[i], i = -n .. 2n
    for (int i = lane; i <= n; i += 32) ra[i] = (int16_t)(vertical ? TOP(i) : LEFT(i));
    if (angle < 0) {
      const int last = (n * angle) >> 5;
      if (last < -1) {
        const int inv = kInvAngle[mode - 11];
        for (int i = -1 - lane; i >= last; i -= 32) {
          const int k = (i * inv + 128) >> 8;
          ra[i] = (int16_t)(vertical ? LEFT(k) : TOP(k));
        }
      }
    } else {
      for (int i = n + 1 + lane; i <= 2 * n; i += 32) ra[i] = (int16_t)(vertical ? TOP(i) : LEFT(i));
    }
    __syncwarp();
  }
  const bool edge = cidx == 0 && n < 32;
  for (int p = lane; p < n * n; p += 32) {
    const int x = p & (n - 1), y = p >> log2;
    int v;
    if (mode == 0) {  // 8.4.4.2.4
      v = ((n - 1 - x) * LEFT(1 + y) + (x + 1) * TOP(1 + n) + (n - 1 - y) * TOP(1 + x) + (y + 1) * LEFT(1 + n) + n) >> (log2 + 1);
    } else if (mode == 1) {  // 8.4.4.2.5
      v = dc;
      if (edge) {
        if (x == 0 && y == 0) v = (LEFT(1) + 2 * dc + TOP(1) + 2) >> 2;
        else if (y == 0) v = (TOP(1 + x) + 3 * dc + 2) >> 2;
        else if (x == 0) v = (LEFT(1 + y) + 3 * dc + 2) >> 2;
      }
    } else {  // 8.4.4.2.6
      const int j = vertical ? y : x, i = vertical ? x : y;
      const int t = (j + 1) * angle, idx = t >> 5, fact = t & 31;
      const int16_t* ra = c.refa;
      v = fact ? ((32 - fact) * ra[i + idx + 1] + fact * ra[i + idx + 2] + 16) >> 5 : ra[i + idx + 1];
      if (edge) {
        if (mode == 26 && x == 0) v = clip8(TOP(1) + ((LEFT(1 + y) - LEFT(0)) >> 1));
        else if (mode == 10 && y == 0) v = clip8(LEFT(1) + ((TOP(1 + x) - TOP(0)) >> 1));
      }
    }
    if (cbf) v = clip8(v + (int)resid[p]);
    buf[(by + y + 1) * stride + bx + x + 16] = (uint8_t)v;
  }
#undef LEFT
#undef TOP
  __syncwarp();
}

struct Shared {
  int progress[kMaxRows];
};

__global__ void __launch_bounds__(256) intra_kernel(Arenas A, int n_slots, int warp_bytes, int log2_ctb_alloc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared* sh = reinterpret_cast<Shared*>(smem_raw);
  const uint32_t tile = blockIdx.x;
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;
  if (A.status[tile].code != 0) return;
  for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) sh->progress[i] = 0;
  __syncthreads();
  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* progress = sh->progress;

  Ctu c;
  c.w = pp->w;
  c.h = pp->h;
  c.ctb = 1 << pp->log2_ctb;
  c.ctb4 = c.ctb >> 2;
  c.lane = lane;
  c.strong_flag = pp->strong_intra_smoothing;
  const int n_planes = pp->chroma ? 3 : 1;
  {
    // carve this warp's scratch: planes (allocated for the batch's largest CTB), ref arrays, mask
    unsigned char* p = smem_raw + sizeof(Shared) + (size_t)slot * warp_bytes;
    const int ctb_a = 1 << log2_ctb_alloc;
    const int sy = 2 * ctb_a + 20, sc = ctb_a + 20;
    c.buf[0] = p;
    c.stride[0] = 2 * c.ctb + 20;
    p += (size_t)(ctb_a + 1) * sy;
    c.buf[1] = p;
    c.stride[1] = c.stride[2] = c.ctb + 20;
    p += (size_t)(ctb_a / 2 + 1) * sc;
    c.buf[2] = p;
    p += (size_t)(ctb_a / 2 + 1) * sc;
    p = reinterpret_cast<unsigned char*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    c.ref = p;
    p += 144;
    c.reff = p;
    p += 144;
    c.refa = reinterpret_cast<int16_t*>(p) + 32;
    p += 2 * 100;
    p = reinterpret_cast<unsigned char*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    c.dmask = reinterpret_cast<uint32_t*>(p);
  }
  const int wctb = pp->wctb, hctb = pp->hctb;
  const int n4sq = c.ctb4 * c.ctb4;
  const uint32_t* tu_map = A.tu_map + tp->tu_off;
  uint8_t* plane[3] = {A.recon + tp->plane_off[0], A.recon + tp->plane_off[1], A.recon + tp->plane_off[2]};
  const int pitch[3] = {pp->pitch_y, pp->pitch_c, pp->pitch_c};

  for (int ry = slot; ry < hctb; ry += n_slots) {
    for (int rx = 0; rx < wctb; rx++) {
      if (ry > 0) {
        const int need = min(rx + 2, wctb);
        if (lane == 0)
          while (progress[ry - 1] < need) __nanosleep(64);
        __syncwarp();
        __threadfence_block();
      }
      c.x_ctb = rx * c.ctb;
      c.y_ctb = ry * c.ctb;
      // ---- apron: left column from the CTU just finished (still in shared memory), top row from HBM ----
      for (int pl = 0; pl < n_planes; pl++) {
        const int sub = pl ? 1 : 0, cs = c.ctb >> sub, st = c.stride[pl];
        uint8_t* b = c.buf[pl];
        if (rx > 0)
          for (int y = lane; y < cs; y += 32) b[(y + 1) * st + 15] = b[(y + 1) * st + 16 + cs - 1];
        if (ry > 0) {
          const int wp = c.w >> sub, x0 = (c.x_ctb >> sub) - 1, y0 = (c.y_ctb >> sub) - 1;
          const uint8_t* src = plane[pl] + (size_t)y0 * pitch[pl];
          for (int i = lane; i <= 2 * cs; i += 32) {
            const int x = x0 + i;
            if (x >= 0 && x < wp) b[15 + i] = __ldcg(src + x);
          }
        }
      }
      if (lane < 8) c.dmask[lane] = 0;
      __syncwarp();
      // ---- transform units of this CTB in z-order -----------------------------------------------------
      const uint32_t ctb_addr = (uint32_t)(ry * wctb + rx);
      const uint32_t* tu = tu_map + (size_t)ctb_addr * n4sq;
      const int16_t* cf0 = A.coeff + tp->coeff_off[0] + (size_t)ctb_addr * n4sq * 16;
      const int16_t* cf1 = A.coeff + tp->coeff_off[1] + (size_t)ctb_addr * (n4sq >> 2) * 16;
      const int16_t* cf2 = A.coeff + tp->coeff_off[2] + (size_t)ctb_addr * (n4sq >> 2) * 16;
      int idx = 0;
      while (idx < n4sq) {
        const uint32_t w = tu[idx];
        if (!(w & TU_ORIGIN)) {  // quadrant outside the picture
          idx++;
          continue;
        }
        const int log2 = (int)tu_log2(w), n = 1 << log2;
        const int ux = (int)compact1((uint32_t)idx), uy = (int)compact1((uint32_t)idx >> 1);
        const int bx = ux << 2, by = uy << 2;
        predict_block(c, 0, bx, by, log2, (int)tu_luma_mode(w), (w & TU_CBF_Y) != 0, cf0 + (size_t)idx * 16);
        {
          const int b4 = n >> 2;
          if (lane < b4) {
            const int bit = (uy + lane) * c.ctb4 + ux;
            atomicOr(&c.dmask[bit >> 5], ((b4 == 32 ? 0xffffffffu : ((1u << b4) - 1u))) << (bit & 31));
          }
          __syncwarp();
        }
        if ((w & TU_HAS_CHROMA) && n_planes == 3) {
          const int log2c = log2 > 2 ? log2 - 1 : 2;
          const int cbx = (log2 > 2 ? bx : bx - 4) >> 1, cby = (log2 > 2 ? by : by - 4) >> 1;
          const size_t off = (size_t)(idx >> 2) * 16;
          const int cm = (int)tu_chroma_mode(w);
          predict_block(c, 1, cbx, cby, log2c, cm, (w & TU_CBF_CB) != 0, cf1 + off);
          predict_block(c, 2, cbx, cby, log2c, cm, (w & TU_CBF_CR) != 0, cf2 + off);
        }
        idx += (n >> 2) * (n >> 2);
      }
      // ---- CTU -> HBM, 4 bytes per lane, rows of the CTU contiguous across lanes -------------------------
      for (int pl = 0; pl < n_planes; pl++) {
        const int sub = pl ? 1 : 0, cs = c.ctb >> sub, st = c.stride[pl];
        const int xo = c.x_ctb >> sub, yo = c.y_ctb >> sub;
        const int wv = min(cs, (c.w >> sub) - xo), hv = min(cs, (c.h >> sub) - yo);
        const int words = cs >> 2;
        const uint8_t* b = c.buf[pl];
        for (int i = lane; i < words * hv; i += 32) {
          const int y = i / words, xw = (i % words) << 2;
          if (xw < wv)
            *reinterpret_cast<uint32_t*>(plane[pl] + (size_t)(yo + y) * pitch[pl] + xo + xw) =
                *reinterpret_cast<const uint32_t*>(b + (y + 1) * st + 16 + xw);
        }
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) progress[ry] = rx + 1;
    }
  }
}

int intra_warp_bytes(int log2_ctb) {
  const int ctb = 1 << log2_ctb;
  size_t b = (size_t)(ctb + 1) * (2 * ctb + 20) + (size_t)2 * (ctb / 2 + 1) * (ctb + 20);
  b = (b + 15) & ~(size_t)15;
  b += 144 + 144 + 200;
  b = (b + 15) & ~(size_t)15;
  b += 32;
  return (int)((b + 15) & ~(size_t)15);
}

}  // namespace

cudaError_t launch_intra(const Arenas& A, int max_log2_ctb, int n_slots, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  const int wb = intra_warp_bytes(max_log2_ctb);
  const size_t smem = sizeof(Shared) + (size_t)n_slots * wb;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(intra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  intra_kernel<<<A.n_tiles, 32 * n_slots, smem, stream>>>(A, n_slots, wb, max_log2_ctb);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
