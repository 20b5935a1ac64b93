// Stage 0 (optional): emulation-prevention removal on the GPU, so the host can ship raw NAL payloads.
// Restates RbspReader::remove_emulation_prevention (hevc/rbsp_reader.rs:11-39): the 0x03 of a 00 00 03 triple is
// dropped only when the following byte is <= 0x03 or the triple ends the payload.  The sequential scan of the
// reference is equivalent to a local test per byte (a dropped 0x03 is never one of the two zeros of another triple),
// which makes it a flag + prefix-sum compaction.  Entry points (7.4.7.1 counts them in raw bytes, the reference never
// converts them: SURVEY Appendix B #2) are re-based from the positions of the removed bytes.
#include <cuda_runtime.h>

#include "kernels.h"

namespace heic {
namespace dev {

namespace {

constexpr int kThreads = 256, kBytesPerThread = 16, kMaxEpb = 1024;
constexpr int kErrTooManyEpb = -3;  // HEIC_E_BITSTREAM

__global__ void __launch_bounds__(kThreads) unescape_kernel(Arenas A, TileParams* tiles, uint32_t* substreams) {
  __shared__ uint32_t epb_pos[kMaxEpb];
  __shared__ uint32_t warp_cnt[kThreads / 32];
  __shared__ uint32_t carry_sm;
  TileParams* tp = tiles + blockIdx.x;
  if (!tp->escaped) return;
  const uint8_t* src = A.raw + tp->raw_off;
  uint8_t* dst = const_cast<uint8_t*>(A.bitstream) + tp->bs_off;
  const uint32_t n = tp->bs_len;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_sm = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += kThreads * kBytesPerThread) {
    const uint32_t p0 = base + tid * kBytesPerThread;
    // bytes p0 - 2 .. p0 + 16 (the raw arena is padded, positions >= n are ignored by the tests below)
    uint8_t b[kBytesPerThread + 3];
#pragma unroll
    for (int j = 0; j < kBytesPerThread + 3; j++) {
      const int64_t p = (int64_t)p0 + j - 2;
      b[j] = (p >= 0 && p < (int64_t)n) ? src[p] : (uint8_t)0xff;
    }
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < kBytesPerThread; j++) {
      const uint32_t p = p0 + j;
      const bool epb = p < n && p >= 2 && b[j + 2] == 3 && b[j + 1] == 0 && b[j] == 0 && (p + 1 >= n || b[j + 3] <= 3);
      mask |= (epb ? 1u : 0u) << j;
    }
    // exclusive scan of the per-thread counts over the CTA
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_cnt[warp] = incl;
    __syncthreads();
    uint32_t before = carry_sm;
    for (int w = 0; w < warp; w++) before += warp_cnt[w];
    uint32_t run = before + incl - cnt;
#pragma unroll
    for (int j = 0; j < kBytesPerThread; j++) {
      const uint32_t p = p0 + j;
      if (p >= n) break;
      if ((mask >> j) & 1u) {
        if (run < kMaxEpb) epb_pos[run] = p;
        run++;
      } else {
        dst[p - run] = b[j + 2];
      }
    }
    __syncthreads();
    if (tid == kThreads - 1) carry_sm = run;  // the last thread's running count is the total so far
    __syncthreads();
  }
  const uint32_t total = carry_sm;
  if (total > kMaxEpb) {
    if (tid == 0) A.status[blockIdx.x].code = kErrTooManyEpb;
    return;
  }
  // removed bytes before raw position x
  auto removed_before = [&](uint32_t x) {
    uint32_t lo = 0, hi = total;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (epb_pos[mid] < x) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  const uint32_t raw_data_off = tp->data_off, n_sub = tp->n_sub, sub_first = tp->sub_first;
  const uint32_t data_off = raw_data_off - removed_before(raw_data_off);
  uint32_t mine = 0;
  if ((uint32_t)tid < n_sub) {
    const uint32_t boundary = raw_data_off + substreams[sub_first + tid];
    mine = boundary - removed_before(boundary) - data_off;
  }
  __syncthreads();
  for (uint32_t k = tid; k < n_sub; k += kThreads) {
    if (k >= (uint32_t)kThreads) {  // more substreams than threads: one at a time (raw values are still in place for k >= kThreads)
      const uint32_t boundary = raw_data_off + substreams[sub_first + k];
      mine = boundary - removed_before(boundary) - data_off;
    }
    substreams[sub_first + k] = mine;
  }
  if (tid == 0) {
    // zero the tail the compaction freed (the CABAC engine prefetches a few bytes past the end)
    for (uint32_t i = n - total; i < n; i++) dst[i] = 0;
    tp->bs_len = n - total;
    tp->data_off = data_off;
    tp->escaped = 0;
  }
}

}  // namespace

cudaError_t launch_unescape(const Arenas& A, TileParams* tiles, uint32_t* substreams, cudaStream_t stream) {
  if (!A.n_tiles) return cudaSuccess;
  unescape_kernel<<<A.n_tiles, kThreads, 0, stream>>>(A, tiles, substreams);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
