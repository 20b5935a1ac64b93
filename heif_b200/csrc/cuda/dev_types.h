// Plain-data structures shared by the host batch builder and the sm_100a kernels.
// Everything here is derived on the host from heic_sps / heic_pps / heic_slice_header
// (include/heic_b200.h) — the kernels never see the raw parameter sets.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HEIC_HD __host__ __device__ __forceinline__
#else
#define HEIC_HD inline
#endif

namespace heic {
namespace dev {

// Flat CABAC context layout (one byte per context: pStateIdx << 1 | valMps).  The reference keys a
// HashMap by (table, idx) (src/cabac/arithmetic.rs:13-20, src/cabac/syntax_element.rs:246-299); the
// offsets below are the same tables concatenated.
enum {
  CTX_SAO_MERGE = 0,        // Table 9-5
  CTX_SAO_TYPE = 1,         // Table 9-6
  CTX_SPLIT_CU = 2,         // Table 9-7, 3
  CTX_CU_TQ_BYPASS = 5,     // Table 9-8
  CTX_PART_MODE = 6,        // Table 9-11
  CTX_PREV_INTRA = 7,       // Table 9-12
  CTX_CHROMA_PRED = 8,      // Table 9-13
  CTX_SPLIT_TRANSFORM = 9,  // Table 9-20, 3
  CTX_CBF_LUMA = 12,        // Table 9-21, 2
  CTX_CBF_CHROMA = 14,      // Table 9-22, 4
  CTX_CU_QP_DELTA = 18,     // Table 9-24, 2
  CTX_TSKIP = 20,           // Table 9-25, luma + chroma
  CTX_LAST_X = 22,          // Table 9-26, 18
  CTX_LAST_Y = 40,          // Table 9-27, 18
  CTX_CSBF = 58,            // Table 9-28, 4
  CTX_SIG = 62,             // Table 9-29, 42
  CTX_GT1 = 104,            // Table 9-30, 24
  CTX_GT2 = 128,            // Table 9-31, 6
  NUM_CTX = 134,
  NUM_CTX_PAD = 136
};

// tu_map word, one per 4x4 luma block in CTB-major z-order; non-zero only at a transform-unit origin.
//   bit 0      TU origin
//   bits 1-2   log2TrafoSize - 2
//   bit 3/4/5  cbf_luma / cbf_cb / cbf_cr (chroma cbfs valid only when bit 6 is set)
//   bit 6      this TU carries the chroma blocks (log2 > 2, or blkIdx 3 of a split 8x8)
//   bits 7-12  IntraPredModeY, bits 13-18 IntraPredModeC, bits 19-24 QpY
//   bits 25-27 transform_skip_flag for Y / Cb / Cr
enum {
  TU_ORIGIN = 1u,
  TU_CBF_Y = 1u << 3,
  TU_CBF_CB = 1u << 4,
  TU_CBF_CR = 1u << 5,
  TU_HAS_CHROMA = 1u << 6,
};
HEIC_HD uint32_t tu_log2(uint32_t w) { return ((w >> 1) & 3u) + 2u; }
HEIC_HD uint32_t tu_luma_mode(uint32_t w) { return (w >> 7) & 63u; }
HEIC_HD uint32_t tu_chroma_mode(uint32_t w) { return (w >> 13) & 63u; }
HEIC_HD uint32_t tu_qp(uint32_t w) { return (w >> 19) & 63u; }
HEIC_HD uint32_t tu_tskip(uint32_t w, int c) { return (w >> (25 + c)) & 1u; }

// Per-image constants (all tiles of a grid share SPS/PPS).
struct PicParams {
  int32_t w, h;                 // luma samples
  int32_t log2_ctb, log2_min_cb, log2_min_tb, log2_max_tb;
  int32_t wctb, hctb;
  int32_t chroma;               // 1: 4:2:0, 0: monochrome
  int32_t max_trafo_depth_intra;
  int32_t cu_qp_delta_enabled, log2_min_cu_qp_delta_size;
  int32_t pps_cb_qp_offset, pps_cr_qp_offset;
  int32_t sign_hiding, tskip_enabled, wpp, strong_intra_smoothing, scaling_enabled;
  int32_t scaling_set;          // index of the ScalingFactor table set (dev::ScalingSet) for this image
  // derived buffer geometry (identical for every tile of the image)
  int32_t w4, h4, w8, h8;       // 4x4 / 8x8 map dims, whole CTBs
  int32_t pitch_y, pitch_c;     // plane row pitches in bytes
  int32_t n_tu;                 // tu_map words per tile = wctb*hctb*(ctb/4)^2
};

// ScalingFactor m[x][y] of 7.4.5, raster y*n+x, per sizeId (4x4 .. 32x32) and colour component.
struct ScalingSet {
  uint8_t f4[3][16];
  uint8_t f8[3][64];
  uint8_t f16[3][256];
  uint8_t f32[3][1024];
};

// Per coded picture (one HEIF grid tile or one single-item image).
struct TileParams {
  uint32_t pic;                 // index into PicParams[]
  uint32_t image, tile_in_image;
  uint32_t bs_off, bs_len;      // un-escaped slice RBSP in the bitstream arena
  uint32_t data_off;            // slice_segment_data() start, relative to bs_off
  // escaped != 0: the host shipped the raw NAL payload (emulation prevention bytes in place) to raw_off in the raw
  // arena; bs_len, data_off and the tile's substream offsets count raw bytes until unescape_kernel has rewritten them
  uint32_t escaped;
  uint64_t raw_off;
  uint32_t sub_first, n_sub;    // substream start offsets (relative to data_off) in the substream array
  int32_t slice_qp;
  int32_t slice_cb_qp_offset, slice_cr_qp_offset;
  int32_t sao_luma, sao_chroma;
  int32_t deblock_disabled, beta_offset_div2, tc_offset_div2;
  // element offsets of this tile's slices of the per-batch arenas
  uint64_t tu_off;              // uint32 words
  uint64_t coeff_off[3];        // int16 elements
  uint64_t plane_off[3];        // bytes (same offsets in the recon and the final arenas)
  uint64_t map4_off;            // bytes: ipm map (w4*h4)
  uint64_t map8_off;            // bytes: ct_depth and qp maps (w8*h8 each)
  uint64_t sao_off;             // uint32 words, 4 per CTB
  uint64_t wpp_off;             // bytes: WPP context snapshots, NUM_CTX_PAD per CTB row
};

// Transform classes: one list (and one launch) per transform size; 4x4 luma (DST) and chroma (DCT) apart.
enum { LIST_32 = 0, LIST_16 = 1, LIST_8 = 2, LIST_4Y = 3, LIST_4C = 4, LIST_CLASSES = 5 };
enum { LIST_COUNT_STRIDE = 32 };  // words between the per-class counters: one cache line each, so their atomics do not serialise
struct uint2_t {
  uint32_t x, y;
};

struct TileStatusDev {
  int32_t code;
  uint32_t bins, ctus, reserved;
};

// Arena base pointers of a resident batch.
struct Arenas {
  const uint8_t* bitstream;
  const uint8_t* raw;            // raw NAL payloads of the tiles with TileParams::escaped
  const uint32_t* substreams;
  const PicParams* pics;
  const TileParams* tiles;
  const ScalingSet* scaling;
  uint32_t* tu_map;
  int16_t* coeff;
  uint8_t* recon;               // reconstruction, deblocked in place
  uint8_t* final_;              // after SAO
  uint8_t* ipm;
  uint8_t* ct_depth;
  uint8_t* qp_map;
  uint32_t* sao;
  uint8_t* wpp_save;
  TileStatusDev* status;
  uint32_t n_tiles;
  // coded transform blocks, one dense list per transform class (built by the transform stage from tu_map)
  uint2_t* tu_list;              // {tile, tu_map entry | colour component << 30}
  uint32_t* list_count;          // LIST_CLASSES counters, LIST_COUNT_STRIDE words apart
  uint32_t list_off[5];          // first element of each class in tu_list
};

}  // namespace dev
}  // namespace heic
