// CABAC kernel: slice data -> syntax (tu_map, TransCoeffLevel, QpY map, SAO parameters).
//
// Replaces the reference's serial tile loop + SliceSegmentReader::read_data (src/heic/decoder.rs:114-119,
// src/hevc/slice.rs:206-231) and its CABAC stack (src/cabac/*).  The arithmetic decoder is inherently
// serial per substream, so the parallelism is substreams = tiles x WPP rows x images:
//
//   TILES == 1   one warp per WPP row of one tile, lane 0 decodes ("warp per substream")
//   TILES == 32  lane l of warp r decodes row r (+ k*R) of the CTA's l-th tile ("thread per substream");
//                the 32 context tables of a warp are interleaved so equal context indices share a 32-byte
//                row of shared memory, and the rows of a tile advance as a wavefront between the warps.
//
// R row slots per CTA share the rows of a tile round-robin; with the 2-CTU WPP lag a 16-CTB-wide tile has
// at most 8 rows in flight, so R = 8 keeps every slot busy.  Rows synchronise through progress counters
// in shared memory (all rows of a tile live in one CTA, hence on one SM); the per-row context snapshots go
// through HBM/L2 so that shared memory only holds the live context tables.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "cabac_parse.cuh"
#include "kernels.h"

namespace heic {
namespace dev {

#if !defined(__CUDA_ARCH__)
// cabac_parse.cuh declares the kernels' dynamic shared memory for the device pass only (the header also builds with g++)
extern __shared__ __align__(16) unsigned char heic_cabac_smem[];
#endif

namespace {

constexpr int kMaxRows = 512;
#ifndef HEIC_CABAC_MIN_CTAS
#define HEIC_CABAC_MIN_CTAS 3  // register budget 80 per thread; 4 (64 regs, more spills) measured equal at batch 592, slower at 296
#endif

struct CtaShared {
  CabacTabs tabs;
  alignas(16) Arenas arenas;  // at kSmemArenasOff: the parser reaches the tile's buffers through it (Parser::arenas())
  int progress[kMaxRows];  // CTUs finished per CTB row (TILES == 32: by all lanes of the row's warp), plus the
                           // group's base (see cabac_kernel): values only ever grow
  int aborted[32];         // per tile of the CTA: use + 1 of the group in which the tile failed
  unsigned group_slot[4];  // persistent CTAs: group taken for use k in slot k & 3, tagged with the use
};

static_assert(offsetof(CtaShared, arenas) == kSmemArenasOff, "Parser::arenas() expects the arena pointers right behind the tables");
constexpr size_t kCtaSharedBytes = (sizeof(CtaShared) + 15) & ~(size_t)15;
// per-thread cold words of the parser (tile index, QP state): CW_COUNT words, one column per thread that parses
template <int TILES>
struct ColdBytes {
  static constexpr size_t value = (size_t)Parser<TILES>::CW_COUNT * (TILES == 32 ? 1024 : 64);
};

// The wavefront hand-shake of parse_rows.  Everything it touches is reached from the kernel's shared-memory symbol and
// the arena pointers kept there -- no pointer members: the structure is live for the whole kernel, and the CABAC kernel
// is bound by its register file.
template <int TILES>
struct SmemSync {
  static constexpr int kSaveStride = 1;
  uint32_t lane, tile;
  int base;  // added to every progress value of the current group (persistent CTAs)
  int use1;  // current use + 1

  __device__ __forceinline__ CtaShared* sh() const { return reinterpret_cast<CtaShared*>(heic_cabac_smem); }
  __device__ __forceinline__ volatile int* progress() const { return sh()->progress; }
  __device__ __forceinline__ volatile int* aborted() const { return &sh()->aborted[TILES == 1 ? 0 : lane]; }

  __device__ __forceinline__ bool wait(int row, int n) {
    unsigned ns = 32;
    volatile int* pr = progress();
    if (TILES == 1) {
      while (pr[row] < base + n && *aborted() != use1) {
        __nanosleep(ns);
        if (ns < 2048) ns *= 2;
      }
    } else {
      // The whole warp polls (one broadcast load) and yields together.  Measured on this pool's B200s a NANOSLEEP returns
      // after ~20 ns whatever its argument (10.9 G of them for 200 warp-seconds of waiting), so the back-off is a run of
      // them per progress probe, not a growing argument.  The poll instructions are not what bounds the kernel: halving
      // them changed nothing, and an mbarrier-suspended wait measured slower.
      __syncwarp();
      while (pr[row] < base + n) {
#pragma unroll
        for (int i = 0; i < 16; i++) __nanosleep(2048);
      }
      __syncwarp();
    }
    __threadfence_block();
    return *aborted() != use1;
  }
  __device__ __forceinline__ void publish(int row, int n) {
    __threadfence_block();
    if (TILES == 1) {
      progress()[row] = base + n;
    } else {
      __syncwarp();
      if (lane == 0) progress()[row] = base + n;
    }
  }
  // this tile's WPP context snapshots in HBM, NUM_CTX_PAD bytes per CTB row (written once and read once per row, so they
  // need not occupy shared memory: that is what bounds occupancy)
  __device__ __forceinline__ uint8_t* save_area(int row) {
    const Arenas& A = sh()->arenas;
    return A.wpp_save + A.tiles[tile].wpp_off + (size_t)row * NUM_CTX_PAD;
  }
  __device__ __forceinline__ void abort(int code) {
    if (code != -100) atomicCAS(&sh()->arenas.status[tile].code, 0, code);
    *aborted() = use1;
    __threadfence_block();
  }
};

}  // namespace

template <int TILES>
__global__ void __launch_bounds__(TILES == 32 ? 256 : 512, TILES == 32 ? HEIC_CABAC_MIN_CTAS : 1) cabac_kernel(Arenas A, const CabacTabs* __restrict__ gtabs,
                                                    const uint32_t* __restrict__ order, int n_slots, uint32_t n_groups,
                                                    uint32_t* group_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaShared* sh = reinterpret_cast<CtaShared*>(smem_raw);
  uint8_t* ctx_all = smem_raw + kCtaSharedBytes + ColdBytes<TILES>::value;

  {  // tables -> shared memory
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gtabs);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh->tabs);
    for (int i = threadIdx.x; i < (int)(sizeof(CabacTabs) / 4); i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) sh->progress[i] = 0;
    if (threadIdx.x < 32) sh->aborted[threadIdx.x] = 0;
    if (threadIdx.x < 4) sh->group_slot[threadIdx.x] = 0;
    if (threadIdx.x == 0) sh->arenas = A;
  }
  __syncthreads();

  const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (TILES == 1 && lane != 0) return;
  // Persistent CTAs (group_counter != nullptr): a CTA takes group after group from a global counter, and its warps
  // move on to the next group one by one, without a CTA-wide barrier — warp 0 starts the first CTB rows of the next
  // group while the last warps still finish the last rows of the current one, so the ramp-up and drain of the WPP
  // wavefront (14 of a 512x512 tile's 46 steps) overlap.  The first warp to need use k claims the group for it.
  for (int use = 0;; use++) {
  uint32_t group = blockIdx.x;
  if (group_counter) {
    uint32_t g = 0;
    if (lane == 0) {
      volatile unsigned* gs = &sh->group_slot[use & 3];
      const unsigned tag = ((unsigned)(use + 1) & 0x7ffu) << 20;
      for (;;) {
        const unsigned e = *gs;
        if ((e & 0x7ff00000u) == tag) {
          if ((e & 0xfffffu) != 0xfffffu) {  // taken and filled in
            g = e & 0xfffffu;
            break;
          }
          __nanosleep(64);
        } else if (atomicCAS(const_cast<unsigned*>(gs), e, tag | 0xfffffu) == e) {
          g = min(atomicAdd(group_counter, 1u), 0xffffeu);
          *gs = tag | g;
          __threadfence_block();
          break;
        }
      }
    }
    group = __shfl_sync(0xffffffffu, g, 0);
    if (group >= n_groups) break;
  } else if (use > 0) {
    break;
  }
  // Idle lanes of a partially filled CTA shadow the geometry of the CTA's first tile but decode nothing.
  const uint32_t my = order[group * TILES + (TILES == 1 ? 0 : lane)];
  const bool active = my != 0xffffffffu;
  const uint32_t tile = active ? my : order[group * TILES];
  const TileParams* tp = A.tiles + tile;
  const PicParams* pp = A.pics + tp->pic;

  Parser<TILES> P;
  P.sm = (uint32_t)__cvta_generic_to_shared(smem_raw);
  P.ctx_off = P.sm + (uint32_t)(kCtaSharedBytes + ColdBytes<TILES>::value) + (uint32_t)slot * NUM_CTX_PAD * TILES + (TILES == 1 ? 0u : (uint32_t)lane);
  P.cold_off = P.sm + (uint32_t)kCtaSharedBytes + (TILES == 32 ? threadIdx.x * 4u : (uint32_t)slot * 4u);
  // keep the addresses in registers: left alone, ptxas rematerialises them from %tid / SR_CgaCtaId (S2R + ALU ops) at every use
  asm volatile("" : "+r"(P.sm), "+r"(P.ctx_off), "+r"(P.cold_off));
  P.cold(Parser<TILES>::CW_TP_LO) = (int)(uint32_t)(unsigned long long)tp;
  P.cold(Parser<TILES>::CW_TP_HI) = (int)(uint32_t)((unsigned long long)tp >> 32);
  P.cold(Parser<TILES>::CW_PP_LO) = (int)(uint32_t)(unsigned long long)pp;
  P.cold(Parser<TILES>::CW_PP_HI) = (int)(uint32_t)((unsigned long long)pp >> 32);
  P.e.data = A.bitstream + tp->bs_off;
  P.err = active ? 0 : -100;

  SmemSync<TILES> sync;
  sync.lane = (uint32_t)lane;
  sync.tile = tile;
  sync.base = use << 12;
  sync.use1 = use + 1;
  if (!active) *sync.aborted() = use + 1;

#if HEIC_CABAC_TRACE  // group timeline for tools/cabac_trace.py: one line per (CTA, group, warp)
  unsigned long long t_begin;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#endif
  const uint32_t ctus = parse_rows<TILES>(P, A.substreams + tp->sub_first, slot, n_slots, sync);
#if HEIC_CABAC_TRACE
  {
    unsigned long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    const unsigned bins_w = __reduce_add_sync(0xffffffffu, active ? P.e.bins : 0u);
    if (lane == 0) printf("CT %u %u %d %d %llu %llu %u %u\n", blockIdx.x, group, use, slot, t_begin, t_end, bins_w, tp->bs_len);
  }
#endif
  if (active) {
    atomicAdd(&A.status[tile].bins, P.e.bins);
    atomicAdd(&A.status[tile].ctus, ctus);
  }
  }  // next group
}

size_t cabac_smem_bytes(int tiles_per_cta, int n_slots) {
  return kCtaSharedBytes + (tiles_per_cta == 32 ? ColdBytes<32>::value : ColdBytes<1>::value) + (size_t)n_slots * NUM_CTX_PAD * tiles_per_cta;
}

cudaError_t launch_cabac(const Arenas& A, const CabacTabs* tabs, const uint32_t* order, uint32_t n_groups,
                         int tiles_per_cta, int n_slots, uint32_t* group_counter, int n_sm, int resident_ctas, cudaStream_t stream) {
  if (!n_groups) return cudaSuccess;
  // Experiment knob: extra (unused) dynamic shared memory per CTA caps the CTAs per SM, leaving registers for
  // kernels of another stream to co-reside with this latency-bound one.
  static const size_t pad = []() { const char* e = getenv("HEIC_B200_CABAC_SMEM_PAD"); return e ? (size_t)atol(e) : (size_t)0; }();
  const size_t smem = cabac_smem_bytes(tiles_per_cta, n_slots) + pad;
  const int threads = 32 * n_slots;
  if (tiles_per_cta == 32) {
    {  // per device, and several devices may be driven from one process: set it on every launch (cheap)
      cudaError_t e = cudaFuncSetAttribute(cabac_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
    }
    // persistent CTAs when there is more than one wave of groups (group ids are 20 bits in the hand-over slot)
    // HEIC_CABAC_MIN_CTAS CTAs of 8 warps are resident per SM: proportionally more when a CTA has fewer row slots
    const uint32_t resident = resident_ctas > 0 ? (uint32_t)resident_ctas : (uint32_t)n_sm * (uint32_t)(HEIC_CABAC_MIN_CTAS * 8 / n_slots);
    if (group_counter && n_groups > resident && n_groups < 0xffff0u)
      cabac_kernel<32><<<resident, threads, smem, stream>>>(A, tabs, order, n_slots, n_groups, group_counter);
    else
      cabac_kernel<32><<<n_groups, threads, smem, stream>>>(A, tabs, order, n_slots, n_groups, nullptr);
  } else {
    cabac_kernel<1><<<n_groups, threads, smem, stream>>>(A, tabs, order, n_slots, n_groups, nullptr);
  }
  return cudaGetLastError();
}

}  // namespace dev
}  // namespace heic
