// Host-side derivation of the kernels' launch parameters from the reference-shaped parameter sets
// (heic_sps mirrors src/hevc/grammar.rs:388-428, heic_pps 511-548, heic_slice_header 551-572; the SPS
// geometry helpers restate grammar.rs:430-508).  No CUDA in this header.
#pragma once
#include <cstring>
#include <string>

#include "../host/hevc_parse.h"
#include "dev_types.h"

namespace heic {

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Throws heic::Error(HEIC_E_UNSUPPORTED / HEIC_E_INVALID_ARG) — the reference's todo!()/unimplemented!()/assert!
// sites on this path (slice.rs:60-63,106-108) become error codes.
inline void make_pic_params(const heic_sps& sps, const heic_pps& pps, dev::PicParams& p) {
  std::memset(&p, 0, sizeof p);
  if (sps.chroma_format_idc > 1) bail(HEIC_E_UNSUPPORTED, "only 4:2:0 and 4:0:0 are supported");
  if (sps.bit_depth_luma_minus8 || sps.bit_depth_chroma_minus8) bail(HEIC_E_UNSUPPORTED, "only 8-bit is supported");
  if (sps.pcm_enabled_flag) bail(HEIC_E_UNSUPPORTED, "PCM is not supported");
  if (pps.transquant_bypass_enabled_flag) bail(HEIC_E_UNSUPPORTED, "cu_transquant_bypass is not supported");
  if (pps.tiles_enabled_flag) bail(HEIC_E_UNSUPPORTED, "HEVC tiles are not supported (HEIF grid tiles are separate pictures)");
  // constrained_intra_pred_flag changes nothing for all-intra pictures.
  p.w = (int)sps.pic_width_in_luma_samples;
  p.h = (int)sps.pic_height_in_luma_samples;
  p.chroma = sps.chroma_format_idc == 1;
  p.log2_min_cb = (int)sps.log2_min_luma_coding_block_size_minus3 + 3;            // grammar.rs MinCbLog2SizeY
  p.log2_ctb = p.log2_min_cb + (int)sps.log2_diff_max_min_luma_coding_block_size;  // CtbLog2SizeY
  p.log2_min_tb = (int)sps.log2_min_luma_transform_block_size_minus2 + 2;
  p.log2_max_tb = p.log2_min_tb + (int)sps.log2_diff_max_min_luma_transform_block_size;
  if (p.w <= 0 || p.h <= 0 || p.w > 16384 || p.h > 16384) bail(HEIC_E_INVALID_ARG, "picture size out of range");
  if (p.log2_ctb < 4 || p.log2_ctb > 6 || p.log2_min_cb < 3 || p.log2_min_cb > p.log2_ctb)
    bail(HEIC_E_BITSTREAM, "coding block sizes out of range");
  if (p.log2_min_tb < 2 || p.log2_max_tb > 5 || p.log2_min_tb >= p.log2_min_cb || p.log2_max_tb > p.log2_ctb)
    bail(HEIC_E_BITSTREAM, "transform block sizes out of range");
  if ((p.w & ((1 << p.log2_min_cb) - 1)) || (p.h & ((1 << p.log2_min_cb) - 1)))
    bail(HEIC_E_BITSTREAM, "picture size is not a multiple of the minimum coding block size");
  const int ctb = 1 << p.log2_ctb;
  p.wctb = (p.w + ctb - 1) >> p.log2_ctb;  // PicWidthInCtbsY
  p.hctb = (p.h + ctb - 1) >> p.log2_ctb;
  if (p.hctb > 512) bail(HEIC_E_UNSUPPORTED, "more than 512 CTB rows");
  p.max_trafo_depth_intra = (int)sps.max_transform_hierarchy_depth_intra;
  p.cu_qp_delta_enabled = (int)pps.cu_qp_delta_enabled_flag;
  p.log2_min_cu_qp_delta_size = p.log2_ctb - (int)pps.diff_cu_qp_delta_depth;
  if (p.cu_qp_delta_enabled && p.log2_min_cu_qp_delta_size < 3) bail(HEIC_E_BITSTREAM, "diff_cu_qp_delta_depth out of range");
  p.pps_cb_qp_offset = pps.pps_cb_qp_offset;
  p.pps_cr_qp_offset = pps.pps_cr_qp_offset;
  p.sign_hiding = (int)pps.sign_data_hiding_enabled_flag;
  p.tskip_enabled = (int)pps.transform_skip_enabled_flag;
  p.wpp = (int)pps.entropy_coding_sync_enabled_flag;
  p.strong_intra_smoothing = (int)sps.strong_intra_smoothing_enabled_flag;
  p.scaling_enabled = (int)sps.scaling_list_enabled_flag;
  p.scaling_set = 0;
  const int ctb4 = ctb >> 2;
  p.w4 = p.wctb * ctb4;
  p.h4 = p.hctb * ctb4;
  p.w8 = p.w4 >> 1;
  p.h8 = p.h4 >> 1;
  p.pitch_y = align_up(p.w, 128);
  p.pitch_c = p.pitch_y >> 1;
  p.n_tu = p.wctb * p.hctb * ctb4 * ctb4;
}

// 7.4.5 ScalingFactor from the (default or transmitted) scaling lists.
inline void build_scaling_set(const heic_sps& sps, const heic_pps& pps, dev::ScalingSet& out) {
  heic_scaling_list def;
  const heic_scaling_list* sl = nullptr;
  if (sps.scaling_list_enabled_flag) {
    if (pps.pps_scaling_list_data_present_flag) sl = &pps.scaling_list;
    else if (sps.sps_scaling_list_data_present_flag) sl = &sps.scaling_list;
    else {
      default_scaling_list(def);
      sl = &def;
    }
  }
  // up-right diagonal scans (6.5.3)
  uint8_t d4[16][2], d8[64][2];
  auto diag = [](int n, uint8_t (*o)[2]) {
    int i = 0, x = 0, y = 0;
    bool stop = false;
    while (!stop) {
      while (y >= 0) {
        if (x < n && y < n) {
          o[i][0] = (uint8_t)x;
          o[i][1] = (uint8_t)y;
          i++;
        }
        y--;
        x++;
      }
      y = x;
      x = 0;
      if (i >= n * n) stop = true;
    }
  };
  diag(4, d4);
  diag(8, d8);
  for (int size_id = 0; size_id < 4; size_id++) {
    const int n = 4 << size_id;
    for (int c = 0; c < 3; c++) {
      uint8_t* f = size_id == 0 ? out.f4[c] : size_id == 1 ? out.f8[c] : size_id == 2 ? out.f16[c] : out.f32[c];
      if (!sl) {
        std::memset(f, 16, (size_t)n * n);
        continue;
      }
      const int matrix_id = (size_id == 3) ? 0 : c;  // 32x32: only the luma list exists for 4:2:0 intra
      const uint8_t* list = sl->list[size_id][matrix_id];
      if (size_id == 0) {
        for (int i = 0; i < 16; i++) f[d4[i][1] * 4 + d4[i][0]] = list[i];
      } else {
        const int rep = n / 8;
        for (int i = 0; i < 64; i++) {
          const int x = d8[i][0], y = d8[i][1];
          for (int j = 0; j < rep; j++)
            for (int k = 0; k < rep; k++) f[(y * rep + j) * n + x * rep + k] = list[i];
        }
        if (size_id >= 2) f[0] = sl->dc[size_id - 2][matrix_id];
      }
    }
  }
}

inline void make_tile_params(const dev::PicParams& pp, const heic_pps& pps, const heic_tile_desc& t,
                             dev::TileParams& o) {
  const heic_slice_header& sh = t.header;
  if (sh.slice_type != 2) bail(HEIC_E_UNSUPPORTED, "only I slices are supported");
  if (!t.rbsp || sh.slice_data_byte_offset > t.rbsp_len) bail(HEIC_E_INVALID_ARG, "slice data offset outside the RBSP");
  if (sh.num_entry_point_offsets > HEIC_MAX_ENTRY_POINTS)  // substream_offset[] has HEIC_MAX_ENTRY_POINTS + 1 entries
    bail(HEIC_E_UNSUPPORTED, "more entry points than heic_slice_header::substream_offset holds");
  if (pp.wpp && (int)sh.num_entry_point_offsets != pp.hctb - 1)
    bail(HEIC_E_UNSUPPORTED, "WPP picture without one entry point per CTB row");
  if (!pp.wpp && sh.num_entry_point_offsets != 0) bail(HEIC_E_UNSUPPORTED, "entry points without WPP (slices/tiles) are not supported");
  o.bs_len = t.rbsp_len;
  o.data_off = sh.slice_data_byte_offset;
  o.n_sub = sh.num_entry_point_offsets + 1;
  o.slice_qp = 26 + pps.init_qp_minus26 + sh.slice_qp_delta;  // cabac/decoder.rs:15
  if (o.slice_qp < 0 || o.slice_qp > 51) bail(HEIC_E_BITSTREAM, "SliceQpY out of range");
  o.slice_cb_qp_offset = sh.slice_cb_qp_offset;
  o.slice_cr_qp_offset = sh.slice_cr_qp_offset;
  o.sao_luma = (int)sh.slice_sao_luma_flag;
  o.sao_chroma = (int)sh.slice_sao_chroma_flag && pp.chroma;
  o.deblock_disabled = (int)sh.slice_deblocking_filter_disabled_flag;
  o.beta_offset_div2 = sh.slice_beta_offset_div2;
  o.tc_offset_div2 = sh.slice_tc_offset_div2;
  for (uint32_t k = 0; k + 1 < o.n_sub; k++)
    if (sh.substream_offset[k + 1] < sh.substream_offset[k] || o.data_off + sh.substream_offset[k + 1] > t.rbsp_len)
      bail(HEIC_E_BITSTREAM, "substream entry point outside the slice data");
}

}  // namespace heic
