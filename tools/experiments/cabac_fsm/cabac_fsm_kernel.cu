// CABAC kernel, state-machine form (cabac_fsm.cuh): slice data -> syntax (tu_map, TransCoeffLevel, QpY map, SAO parameters).
//
// Thread mapping: a CTA has R row slots (warps) x 32 COLUMNS (lanes).  A column decodes one tile at a time: the thread in
// warp r of column l decodes CTB rows r, r + R, ... of the column's current tile, the rows of the tile advancing as the WPP
// wavefront between the R threads of the column.  Everything a column shares lives in shared memory and is PER COLUMN:
//   progress[r][l]   CTUs finished by thread (r, l), as (use, row, count) packed so that the word only ever grows
//   aborted[l]       use + 1 of the tile that failed
//   ring[4][l]       the tile of the column's use k (k & 3), claimed by whichever of its threads needs it first; a thread
//                    that has finished its rows of a tile moves on to the column's next tile on its own
// so the lanes of a warp are independent: none waits for another's tile, row, CTU or bin.  A warp's instruction stream is
// the loop of Fsm::step(): one arithmetic-decoder operation for all lanes, then each lane's state body.
// Tiles come from one queue per launch (sorted by slice size, heaviest first, dealt so that neighbours in the queue
// are different pictures); the first 32 x gridDim entries are assigned statically, the rest claimed with an atomic counter.
//
// Replaces the reference's serial tile loop + SliceSegmentReader::read_data (src/heic/decoder.rs:114-119,
// src/hevc/slice.rs:206-231) and its CABAC stack (src/cabac/*).
#include <cuda_runtime.h>

#include "cabac_fsm.cuh"
#include "../../../heif_b200/csrc/cuda/kernels.h"

namespace heic {
namespace dev {

#if !defined(__CUDA_ARCH__)
// cabac_parse.cuh declares the kernel's dynamic shared memory for the device pass only (the header also builds with g++)
extern __shared__ __align__(16) unsigned char heic_cabac_smem[];
#endif

namespace {

#ifndef HEIC_CABAC_FSM_MIN_CTAS
#define HEIC_CABAC_FSM_MIN_CTAS 3
#endif
constexpr int kMaxThreads = 256;  // 8 row slots x 32 columns

struct FsmShared {
  CabacTabs tabs;  // must stay first: Engine::decision and the scan tables address it at offset 0
  Arenas arenas;
  const uint32_t* order;
  uint32_t* counter;
  uint32_t n_entries, first_dynamic;
  uint32_t progress[8][32];
  uint32_t aborted[32];
  uint32_t ring_use[4][32];      // (use + 1) << 1 | filled
  uint32_t ring_tile[4][32];
  uint32_t ring_readers[4][32];  // threads of the column that have read the entry
  uint32_t cold[CW_COUNT][kMaxThreads];
};

// The column's tile of use `use`: claimed by the first of its threads that asks, read from the ring by the others.  Out of
// line and free-standing (a non-inlined member would force the whole state machine into local memory through `this`).
__device__ __noinline__ uint32_t fsm_acquire_tile(uint32_t use, uint32_t lane, uint32_t slots) {
    FsmShared* S = reinterpret_cast<FsmShared*>(heic_cabac_smem);
    const uint32_t k = use & 3u, tag = (use + 1u) << 1;
    volatile uint32_t* su = &S->ring_use[k][lane];
    const uint32_t ent = *su;
    if ((ent & ~1u) == tag) {
      if (!(ent & 1u)) return TILE_RETRY;  // being filled in
      __threadfence_block();
      const uint32_t t = *(volatile uint32_t*)&S->ring_tile[k][lane];
      atomicAdd(&S->ring_readers[k][lane], 1u);
      return t;
    }
    // the entry still belongs to use - 4: it may be recycled once every thread of the column has read it
    if (use >= 4u && *(volatile uint32_t*)&S->ring_readers[k][lane] < slots) return TILE_RETRY;
    if (atomicCAS(const_cast<uint32_t*>(su), ent, tag) != ent) return TILE_RETRY;
    S->ring_readers[k][lane] = 1u;  // this thread
    uint32_t t = TILE_NONE;
    if (use == 0u) {
      const uint32_t idx = blockIdx.x * 32u + lane;
      if (idx < S->n_entries) t = S->order[idx];
    }
    while (t == TILE_NONE) {  // padding entries of the launch order are skipped
      const uint32_t idx = S->first_dynamic + atomicAdd(S->counter, 1u);
      if (idx >= S->n_entries) break;
      t = S->order[idx];
    }
    S->ring_tile[k][lane] = t;
    __threadfence_block();
    *su = tag | 1u;
    return t;
}

__device__ __noinline__ void fsm_abort_tile(int code, uint32_t tile, uint32_t lane, uint32_t use) {
  FsmShared* S = reinterpret_cast<FsmShared*>(heic_cabac_smem);
  if (code != -100) atomicCAS(&S->arenas.status[tile].code, 0, code);
  *(volatile uint32_t*)&S->aborted[lane] = use + 1u;
  __threadfence_block();
}
__device__ __noinline__ void fsm_finish_tile(uint32_t tile, uint32_t bins, uint32_t ctus) {
  TileStatusDev* st = reinterpret_cast<FsmShared*>(heic_cabac_smem)->arenas.status + tile;
  if (bins) atomicAdd(&st->bins, bins);
  if (ctus) atomicAdd(&st->ctus, ctus);
}

struct FsmEnvDev {
  uint32_t ctx_off;   // byte offset of this thread's context table (contexts 32 bytes apart: one row per context index)
  uint32_t cold_off;  // byte offset of this thread's column of cold words
  uint32_t lane, my_slot, slots, cur_use;

  __device__ __forceinline__ FsmShared* sh() const { return reinterpret_cast<FsmShared*>(heic_cabac_smem); }
  __device__ __forceinline__ const CabacTabs* tabs() const { return reinterpret_cast<const CabacTabs*>(heic_cabac_smem); }
  __device__ __forceinline__ uint32_t ld_ctx(int idx) const { return heic_cabac_smem[ctx_off + idx * 32]; }
  __device__ __forceinline__ void st_ctx(int idx, uint32_t v) { heic_cabac_smem[ctx_off + idx * 32] = (uint8_t)v; }
  __device__ __forceinline__ uint32_t& cw(int j) {
    return *reinterpret_cast<uint32_t*>(heic_cabac_smem + cold_off + j * (kMaxThreads * 4));
  }
  __device__ __forceinline__ const Arenas* arenas() const { return &sh()->arenas; }
  __device__ __forceinline__ int slot() const { return (int)my_slot; }
  __device__ __forceinline__ int n_slots() const { return (int)slots; }
  __device__ __forceinline__ static uint32_t key(uint32_t use, int row, int n) { return ((use * 1024u + (uint32_t)row) << 10) | (uint32_t)n; }

  __device__ __forceinline__ uint32_t acquire_tile(uint32_t use) {
    cur_use = use;
    return fsm_acquire_tile(use, lane, slots);
  }
  __device__ __forceinline__ uint32_t wait_key(int row, int need) const { return key(cur_use, row, need); }
  __device__ __forceinline__ int wait_ready(uint32_t k) const {
    const FsmShared* S = sh();
    const uint32_t up = my_slot == 0u ? slots - 1u : my_slot - 1u;  // the row above belongs to the slot before this thread's
    const uint32_t p = *(volatile const uint32_t*)&S->progress[up][lane];
    if (p >= k) {
      __threadfence_block();
      return 1;
    }
    return *(volatile const uint32_t*)&S->aborted[lane] == cur_use + 1u ? 2 : 0;
  }
  __device__ __forceinline__ void publish(int row, int n) {
    __threadfence_block();
    *(volatile uint32_t*)&sh()->progress[my_slot][lane] = key(cur_use, row, n);
  }
  __device__ __forceinline__ void abort_tile(int code) { fsm_abort_tile(code, cw(CW_TILE), lane, cur_use); }
  __device__ __forceinline__ void finish_tile(uint32_t tile, uint32_t bins, uint32_t ctus) { fsm_finish_tile(tile, bins, ctus); }
};

constexpr size_t kFsmSharedBytes = (sizeof(FsmShared) + 127) & ~(size_t)127;

}  // namespace

__global__ void __launch_bounds__(kMaxThreads, HEIC_CABAC_FSM_MIN_CTAS)
cabac_fsm_kernel(Arenas A, const CabacTabs* __restrict__ gtabs, const uint32_t* __restrict__ order, uint32_t n_entries, int n_slots,
                 uint32_t* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FsmShared* S = reinterpret_cast<FsmShared*>(smem_raw);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gtabs);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S->tabs);
    for (int i = threadIdx.x; i < (int)(sizeof(CabacTabs) / 4); i += blockDim.x) dst[i] = src[i];
    uint32_t* z = &S->progress[0][0];
    const int n_zero = (int)((sizeof(S->progress) + sizeof(S->aborted) + sizeof(S->ring_use) + sizeof(S->ring_tile) + sizeof(S->ring_readers)) / 4);
    for (int i = threadIdx.x; i < n_zero; i += blockDim.x) z[i] = 0;
    if (threadIdx.x == 0) {
      S->arenas = A;
      S->order = order;
      S->counter = counter;
      S->n_entries = n_entries;
      S->first_dynamic = gridDim.x * 32u;
    }
  }
  __syncthreads();

  Fsm<FsmEnvDev> F;
  F.env.lane = threadIdx.x & 31u;
  F.env.my_slot = threadIdx.x >> 5;
  F.env.slots = (uint32_t)n_slots;
  F.env.cur_use = 0;
  F.env.ctx_off = (uint32_t)kFsmSharedBytes + (F.env.my_slot * NUM_CTX_PAD) * 32u + F.env.lane;
  F.env.cold_off = (uint32_t)offsetof(FsmShared, cold) + threadIdx.x * 4u;
  // keep the offsets in registers: left alone, ptxas rematerialises them from %tid at every use
  asm volatile("" : "+r"(F.env.ctx_off), "+r"(F.env.cold_off));
  F.init();
  // Watchdog: no stream, however malformed, may hang the device.  Every state makes progress or idles on another thread
  // of its column, so this never fires by design; if it does the thread's tile is flagged and the thread leaves.
  const long long t_start = clock64();
  uint32_t iter = 0;
  // All 32 lanes stay in the loop until the last one is done, and meet at the vote below in EVERY iteration.  Without that
  // convergence point the lanes that part ways in the state switch would only have to meet again at the loop's exit, and
  // the warp would fall apart into groups that run the whole loop one after the other (measured: 7x slower).
  bool alive = true;
  for (;;) {
    if (alive) alive = F.step();
    const bool busy = alive && !(F.op == OP_NONE && (F.st == S_WAIT || F.st == S_TILE_NEXT));
    const unsigned any_alive = __ballot_sync(0xffffffffu, alive), any_busy = __ballot_sync(0xffffffffu, busy);
    if (!any_alive) break;
    // every lane idle (waiting for the row above, or for its column's next tile): yield the issue slots
    if (!any_busy) __nanosleep(200);
    if ((++iter & 0xfffffu) == 0u && clock64() - t_start > 120000000000ll) {  // ~ one minute
      if (alive) F.env.abort_tile(-5);
      break;
    }
  }
}

size_t cabac_fsm_smem_bytes(int n_slots) { return kFsmSharedBytes + (size_t)n_slots * NUM_CTX_PAD * 32; }

cudaError_t launch_cabac_fsm(const Arenas& A, const CabacTabs* tabs, const uint32_t* order, uint32_t n_entries, uint32_t n_tiles,
                             int n_slots, uint32_t* counter, int n_sm, int resident_ctas, cudaStream_t stream) {
  if (!n_entries || !n_tiles) return cudaSuccess;
  if (n_slots < 1 || n_slots > 8) return cudaErrorInvalidValue;
  const size_t smem = cabac_fsm_smem_bytes(n_slots);
  cudaError_t e = cudaFuncSetAttribute(cabac_fsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cabac_fsm_smem_bytes(8));
  if (e != cudaSuccess) return e;
  // as many CTAs as stay resident (HEIC_CABAC_FSM_MIN_CTAS x 8 warps per SM; proportionally more with fewer row slots), but
  // no more columns than tiles
  const uint32_t resident = resident_ctas > 0 ? (uint32_t)resident_ctas : (uint32_t)n_sm * (uint32_t)(HEIC_CABAC_FSM_MIN_CTAS * 8 / n_slots);
  const uint32_t wanted = (n_tiles + 31u) / 32u;
  const uint32_t grid = wanted < resident ? wanted : resident;
  cabac_fsm_kernel<<<grid, 32 * n_slots, smem, stream>>>(A, tabs, order, n_entries, n_slots, counter);
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
