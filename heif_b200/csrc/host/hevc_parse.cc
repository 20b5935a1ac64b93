// HEVC parameter-set and slice-segment-header parsing on the host.
// Restates the reference's src/hevc/parameter_set_reader.rs and src/hevc/slice.rs:44-204 (same
// syntax order), but keeps what the reference parses-and-drops and the GPU path needs:
// scaling lists (parameter_set_reader.rs:97-103,203-222), video_full_range_flag (:281),
// the slice-data byte offset and EPB-corrected substream offsets.
#include <algorithm>
#include <cstring>

#include "hevc_parse.h"

namespace heic {

namespace {

// 7.3.3 profile_tier_level (parameter_set_reader.rs:493-551 skips it the same way).
void skip_profile_tier_level(RbspReader& r, bool profile_present, uint32_t max_sub_layers_minus1) {
  if (profile_present) {
    r.read_bits(2 + 1 + 5);
    r.read_bits(32);
    r.read_bits(48);
  }
  r.read_bits(8);  // general_level_idc
  bool sub_profile[8] = {false}, sub_level[8] = {false};
  for (uint32_t i = 0; i < max_sub_layers_minus1; ++i) {
    sub_profile[i] = r.read_flag();
    sub_level[i] = r.read_flag();
  }
  if (max_sub_layers_minus1 > 0)
    for (uint32_t i = max_sub_layers_minus1; i < 8; ++i) r.read_bits(2);
  for (uint32_t i = 0; i < max_sub_layers_minus1; ++i) {
    if (sub_profile[i]) {
      r.read_bits(8);
      r.read_bits(32);
      r.read_bits(48);
    }
    if (sub_level[i]) r.read_bits(8);
  }
}

// Default lists, Tables 7-5 / 7-6 (diagonal scan order).  Intra = matrixId 0..2, inter = 3..5.
const uint8_t kDefault8x8Intra[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18, 17, 18, 18, 17, 18, 21,
    19, 20, 21, 20, 19, 21, 24, 22, 22, 24, 24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29,
    31, 35, 35, 31, 29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115};
const uint8_t kDefault8x8Inter[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 20,
    20, 20, 20, 20, 20, 20, 24, 24, 24, 24, 24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28,
    28, 28, 28, 28, 28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91};

void set_default_list(heic_scaling_list& sl, int size_id, int matrix_id) {
  if (size_id == 0) {
    std::memset(sl.list[0][matrix_id], 16, 64);
  } else {
    std::memcpy(sl.list[size_id][matrix_id], matrix_id < 3 ? kDefault8x8Intra : kDefault8x8Inter, 64);
    if (size_id >= 2) sl.dc[size_id - 2][matrix_id] = 16;
  }
}

// 7.3.4 scaling_list_data (the reference's skip_scaling_list_data, parameter_set_reader.rs:203-222,
// walks the same syntax and throws the values away).
void parse_scaling_list_data(RbspReader& r, heic_scaling_list& sl) {
  for (int size_id = 0; size_id < 4; ++size_id) {
    for (int matrix_id = 0; matrix_id < 6; matrix_id += (size_id == 3 ? 3 : 1)) {
      bool pred_mode_flag = r.read_flag();
      if (!pred_mode_flag) {
        uint32_t delta = r.read_ue();
        if (size_id == 3) delta *= 3;
        if (delta > static_cast<uint32_t>(matrix_id)) bail(HEIC_E_BITSTREAM, "scaling_list_pred_matrix_id_delta out of range");
        if (delta == 0) {
          set_default_list(sl, size_id, matrix_id);
        } else {
          int ref = matrix_id - static_cast<int>(delta);
          std::memcpy(sl.list[size_id][matrix_id], sl.list[size_id][ref], 64);
          if (size_id >= 2) sl.dc[size_id - 2][matrix_id] = sl.dc[size_id - 2][ref];
        }
      } else {
        int coef_num = std::min(64, 1 << (4 + (size_id << 1)));
        int next = 8;
        if (size_id > 1) {
          int dc = r.read_se() + 8;
          if (dc < 1 || dc > 255) bail(HEIC_E_BITSTREAM, "scaling_list_dc_coef out of range");
          next = dc;
          sl.dc[size_id - 2][matrix_id] = static_cast<uint8_t>(dc);
        }
        for (int i = 0; i < coef_num; ++i) {
          int delta = r.read_se();
          next = (next + delta + 256) % 256;
          sl.list[size_id][matrix_id][i] = static_cast<uint8_t>(next);
        }
      }
    }
  }
  // 32x32 chroma lists (matrixId 1,2,4,5) only exist for 4:4:4; derive them from 16x16 as RExt does
  // so the table is fully defined.
  for (int m : {1, 2, 4, 5}) {
    std::memcpy(sl.list[3][m], sl.list[2][m], 64);
    sl.dc[1][m] = sl.dc[0][m];
  }
}

void skip_st_ref_pic_set(RbspReader& r, uint32_t idx) {
  bool inter = idx != 0 ? r.read_flag() : false;
  if (inter) bail(HEIC_E_UNSUPPORTED, "inter RPS prediction in SPS (not a still-picture stream)");
  uint32_t neg = r.read_ue(), pos = r.read_ue();
  if (neg > 16 || pos > 16) bail(HEIC_E_BITSTREAM, "st_ref_pic_set too large");
  for (uint32_t i = 0; i < neg + pos; ++i) {
    r.read_ue();
    r.read_flag();
  }
}

void skip_sub_layer_hrd(RbspReader& r, uint32_t cpb_cnt, bool sub_pic) {
  for (uint32_t i = 0; i < cpb_cnt; ++i) {
    r.read_ue();
    r.read_ue();
    if (sub_pic) {
      r.read_ue();
      r.read_ue();
    }
    r.read_flag();
  }
}

// E.2.2 hrd_parameters — parsed fully (the reference stops after three flags,
// parameter_set_reader.rs:338-349, which would desynchronise on a stream that carries HRD).
void skip_hrd_parameters(RbspReader& r, bool common, uint32_t max_sub_layers_minus1) {
  bool nal = false, vcl = false, sub_pic = false;
  if (common) {
    nal = r.read_flag();
    vcl = r.read_flag();
    if (nal || vcl) {
      sub_pic = r.read_flag();
      if (sub_pic) r.read_bits(8 + 5 + 1 + 5);
      r.read_bits(4 + 4);
      if (sub_pic) r.read_bits(4);
      r.read_bits(5 + 5 + 5);
    }
  }
  for (uint32_t i = 0; i <= max_sub_layers_minus1; ++i) {
    bool fixed_general = r.read_flag();
    bool fixed_within = fixed_general ? true : r.read_flag();
    bool low_delay = false;
    if (fixed_within) r.read_ue();
    else low_delay = r.read_flag();
    uint32_t cpb_cnt = 1;
    if (!low_delay) cpb_cnt = r.read_ue() + 1;
    if (cpb_cnt > 32) bail(HEIC_E_BITSTREAM, "cpb_cnt_minus1 out of range");
    if (nal) skip_sub_layer_hrd(r, cpb_cnt, sub_pic);
    if (vcl) skip_sub_layer_hrd(r, cpb_cnt, sub_pic);
  }
}

// E.2.1 vui_parameters (parameter_set_reader.rs:252-336).
void parse_vui(RbspReader& r, heic_sps& sps) {
  if (r.read_flag()) {  // aspect_ratio_info_present_flag
    if (r.read_u8(8) == 255) r.read_bits(32);
  }
  if (r.read_flag()) r.read_flag();  // overscan
  if (r.read_flag()) {               // video_signal_type_present_flag
    r.read_bits(3);
    sps.video_full_range_flag = r.read_flag();
    if (r.read_flag()) {
      sps.colour_primaries = r.read_u8(8);
      sps.transfer_characteristics = r.read_u8(8);
      sps.matrix_coeffs = r.read_u8(8);
    }
  }
  if (r.read_flag()) {  // chroma_loc_info_present_flag
    r.read_ue();
    r.read_ue();
  }
  r.read_bits(3);       // neutral_chroma_indication, field_seq, frame_field_info_present
  if (r.read_flag()) {  // default_display_window_flag
    r.read_ue();
    r.read_ue();
    r.read_ue();
    r.read_ue();
  }
  if (r.read_flag()) {  // vui_timing_info_present_flag
    r.read_bits(32);
    r.read_bits(32);
    if (r.read_flag()) r.read_ue();
    if (r.read_flag()) skip_hrd_parameters(r, true, sps.sps_max_sub_layers_minus1);
  }
  if (r.read_flag()) {  // bitstream_restriction_flag
    r.read_bits(3);
    for (int i = 0; i < 5; ++i) r.read_ue();
  }
}

}  // namespace

void default_scaling_list(heic_scaling_list& sl) {
  for (int s = 0; s < 4; ++s)
    for (int m = 0; m < 6; ++m) set_default_list(sl, s, m);
}

// parameter_set_reader.rs:36-201
heic_sps sequence_parameter_set_rbsp(const uint8_t* data, size_t len) {
  RbspReader r(data, len);
  heic_sps s;
  std::memset(&s, 0, sizeof s);
  s.colour_primaries = s.transfer_characteristics = s.matrix_coeffs = 2;
  default_scaling_list(s.scaling_list);

  s.sps_video_parameter_set_id = r.read_u8(4);
  s.sps_max_sub_layers_minus1 = r.read_u8(3);
  s.sps_temporal_id_nesting_flag = r.read_flag();
  skip_profile_tier_level(r, true, s.sps_max_sub_layers_minus1);
  s.sps_seq_parameter_set_id = r.read_ue();
  s.chroma_format_idc = r.read_ue();
  ensure(s.chroma_format_idc <= 3, HEIC_E_BITSTREAM, "invalid chroma_format_idc");
  if (s.chroma_format_idc == 3) s.separate_colour_plane_flag = r.read_flag();
  s.pic_width_in_luma_samples = r.read_ue();
  s.pic_height_in_luma_samples = r.read_ue();
  s.conformance_window_flag = r.read_flag();
  if (s.conformance_window_flag) {
    s.conf_win_left_offset = r.read_ue();
    s.conf_win_right_offset = r.read_ue();
    s.conf_win_top_offset = r.read_ue();
    s.conf_win_bottom_offset = r.read_ue();
  }
  s.bit_depth_luma_minus8 = r.read_ue();
  s.bit_depth_chroma_minus8 = r.read_ue();
  s.log2_max_pic_order_cnt_lsb_minus4 = r.read_ue();
  ensure(s.log2_max_pic_order_cnt_lsb_minus4 <= 12, HEIC_E_BITSTREAM, "log2_max_pic_order_cnt_lsb_minus4 > 12");
  bool sub_layer_ordering_info_present = r.read_flag();
  for (uint32_t i = sub_layer_ordering_info_present ? 0 : s.sps_max_sub_layers_minus1;
       i <= s.sps_max_sub_layers_minus1; ++i) {
    r.read_ue();
    r.read_ue();
    r.read_ue();
  }
  s.log2_min_luma_coding_block_size_minus3 = r.read_ue();
  s.log2_diff_max_min_luma_coding_block_size = r.read_ue();
  s.log2_min_luma_transform_block_size_minus2 = r.read_ue();
  s.log2_diff_max_min_luma_transform_block_size = r.read_ue();
  s.max_transform_hierarchy_depth_inter = r.read_ue();
  s.max_transform_hierarchy_depth_intra = r.read_ue();
  s.scaling_list_enabled_flag = r.read_flag();
  if (s.scaling_list_enabled_flag) {
    s.sps_scaling_list_data_present_flag = r.read_flag();
    if (s.sps_scaling_list_data_present_flag) parse_scaling_list_data(r, s.scaling_list);
  }
  s.amp_enabled_flag = r.read_flag();
  s.sample_adaptive_offset_enabled_flag = r.read_flag();
  s.pcm_enabled_flag = r.read_flag();
  if (s.pcm_enabled_flag) {
    s.pcm_sample_bit_depth_luma_minus1 = r.read_u8(4);
    s.pcm_sample_bit_depth_chroma_minus1 = r.read_u8(4);
    s.log2_min_pcm_luma_coding_block_size_minus3 = r.read_ue();
    s.log2_diff_max_min_pcm_luma_coding_block_size = r.read_ue();
    s.pcm_loop_filter_disabled_flag = r.read_flag();
  }
  s.num_short_term_ref_pic_sets = r.read_ue();
  ensure(s.num_short_term_ref_pic_sets <= 64, HEIC_E_BITSTREAM, "num_short_term_ref_pic_sets > 64");
  for (uint32_t i = 0; i < s.num_short_term_ref_pic_sets; ++i) skip_st_ref_pic_set(r, i);
  s.long_term_ref_pics_present_flag = r.read_flag();
  if (s.long_term_ref_pics_present_flag) {
    uint32_t n = r.read_ue();
    ensure(n <= 32, HEIC_E_BITSTREAM, "num_long_term_ref_pics_sps > 32");
    for (uint32_t i = 0; i < n; ++i) {
      r.read_bits(s.log2_max_pic_order_cnt_lsb_minus4 + 4);
      r.read_flag();
    }
  }
  s.sps_temporal_mvp_enabled_flag = r.read_flag();
  s.strong_intra_smoothing_enabled_flag = r.read_flag();
  s.vui_parameters_present_flag = r.read_flag();
  if (s.vui_parameters_present_flag) parse_vui(r, s);
  bool sps_extension_present = r.read_flag();
  if (sps_extension_present) {
    // The reference rejects any extension (parameter_set_reader.rs:153-158).  We accept a range
    // extension whose tools are all off (what Apple's 4:0:0 gain-map stream carries).
    bool range_ext = r.read_flag();
    bool multilayer_ext = r.read_flag();
    bool ext_3d = r.read_flag();
    bool scc_ext = r.read_flag();
    uint32_t ext_4bits = r.read_u8(4);
    if (multilayer_ext || ext_3d || scc_ext || ext_4bits)
      bail(HEIC_E_UNSUPPORTED, "SPS multilayer/3D/SCC extension");
    if (range_ext) {
      uint32_t tools = r.read_u32(9);
      if (tools) bail(HEIC_E_UNSUPPORTED, "SPS range-extension coding tools enabled");
    }
  }
  // Geometry constraints (7.4.3.2.1).
  uint32_t log2_min_cb = s.log2_min_luma_coding_block_size_minus3 + 3;
  uint32_t log2_ctb = log2_min_cb + s.log2_diff_max_min_luma_coding_block_size;
  uint32_t log2_min_tb = s.log2_min_luma_transform_block_size_minus2 + 2;
  uint32_t log2_max_tb = log2_min_tb + s.log2_diff_max_min_luma_transform_block_size;
  ensure(log2_ctb >= 4 && log2_ctb <= 6, HEIC_E_BITSTREAM, "CtbLog2SizeY outside 4..6");
  ensure(log2_max_tb <= 5 && log2_max_tb <= log2_ctb && log2_min_tb < log2_min_cb, HEIC_E_BITSTREAM,
         "invalid transform block size range");
  ensure(s.pic_width_in_luma_samples > 0 && s.pic_height_in_luma_samples > 0 &&
             s.pic_width_in_luma_samples % (1u << log2_min_cb) == 0 &&
             s.pic_height_in_luma_samples % (1u << log2_min_cb) == 0,
         HEIC_E_BITSTREAM, "picture size is not a multiple of MinCbSizeY");
  return s;
}

// parameter_set_reader.rs:351-491
heic_pps picture_parameter_set_rbsp(const uint8_t* data, size_t len) {
  RbspReader r(data, len);
  heic_pps p;
  std::memset(&p, 0, sizeof p);
  default_scaling_list(p.scaling_list);
  p.pps_pic_parameter_set_id = r.read_ue();
  p.pps_seq_parameter_set_id = r.read_ue();
  p.dependent_slice_segments_enabled_flag = r.read_flag();
  p.output_flag_present_flag = r.read_flag();
  p.num_extra_slice_header_bits = r.read_u8(3);
  p.sign_data_hiding_enabled_flag = r.read_flag();
  p.cabac_init_present_flag = r.read_flag();
  p.num_ref_idx_l0_default_active_minus1 = r.read_ue();
  p.num_ref_idx_l1_default_active_minus1 = r.read_ue();
  p.init_qp_minus26 = r.read_se();
  p.constrained_intra_pred_flag = r.read_flag();
  p.transform_skip_enabled_flag = r.read_flag();
  p.cu_qp_delta_enabled_flag = r.read_flag();
  if (p.cu_qp_delta_enabled_flag) p.diff_cu_qp_delta_depth = r.read_ue();
  p.pps_cb_qp_offset = r.read_se();
  p.pps_cr_qp_offset = r.read_se();
  p.pps_slice_chroma_qp_offsets_present_flag = r.read_flag();
  p.weighted_pred_flag = r.read_flag();
  p.weighted_bipred_flag = r.read_flag();
  p.transquant_bypass_enabled_flag = r.read_flag();
  p.tiles_enabled_flag = r.read_flag();
  p.entropy_coding_sync_enabled_flag = r.read_flag();
  if (p.tiles_enabled_flag) {
    p.num_tile_columns_minus1 = r.read_ue();
    p.num_tile_rows_minus1 = r.read_ue();
    ensure(p.num_tile_columns_minus1 < 64 && p.num_tile_rows_minus1 < 64, HEIC_E_BITSTREAM, "too many HEVC tiles");
    p.uniform_spacing_flag = r.read_flag();
    if (!p.uniform_spacing_flag) {
      for (uint32_t i = 0; i < p.num_tile_columns_minus1; ++i) r.read_ue();
      for (uint32_t i = 0; i < p.num_tile_rows_minus1; ++i) r.read_ue();
    }
    p.loop_filter_across_tiles_enabled_flag = r.read_flag();
  }
  p.pps_loop_filter_across_slices_enabled_flag = r.read_flag();
  p.deblocking_filter_control_present_flag = r.read_flag();
  if (p.deblocking_filter_control_present_flag) {
    p.deblocking_filter_override_enabled_flag = r.read_flag();
    p.pps_deblocking_filter_disabled_flag = r.read_flag();
    if (!p.pps_deblocking_filter_disabled_flag) {
      p.pps_beta_offset_div2 = r.read_se();
      p.pps_tc_offset_div2 = r.read_se();
    }
  }
  p.pps_scaling_list_data_present_flag = r.read_flag();
  if (p.pps_scaling_list_data_present_flag) parse_scaling_list_data(r, p.scaling_list);
  p.lists_modification_present_flag = r.read_flag();
  p.log2_parallel_merge_level_minus2 = r.read_ue();
  p.slice_segment_header_extension_present_flag = r.read_flag();
  bool pps_extension_present = r.read_flag();
  if (pps_extension_present) {
    bool range_ext = r.read_flag();
    bool multilayer_ext = r.read_flag();
    bool ext_3d = r.read_flag();
    bool scc_ext = r.read_flag();
    uint32_t ext_4bits = r.read_u8(4);
    if (multilayer_ext || ext_3d || scc_ext || ext_4bits) bail(HEIC_E_UNSUPPORTED, "PPS multilayer/3D/SCC extension");
    if (range_ext) {
      // pps_range_extension: only the all-off form is accepted.
      if (p.transform_skip_enabled_flag && r.read_ue() != 0)
        bail(HEIC_E_UNSUPPORTED, "log2_max_transform_skip_block_size_minus2 != 0");
      bool cross_component = r.read_flag();
      bool chroma_qp_offset_list = r.read_flag();
      if (cross_component || chroma_qp_offset_list) bail(HEIC_E_UNSUPPORTED, "PPS range-extension coding tools enabled");
      uint32_t sao_shift_luma = r.read_ue(), sao_shift_chroma = r.read_ue();
      if (sao_shift_luma || sao_shift_chroma) bail(HEIC_E_UNSUPPORTED, "log2_sao_offset_scale != 0");
    }
  }
  ensure(p.init_qp_minus26 >= -26 && p.init_qp_minus26 <= 25, HEIC_E_BITSTREAM, "init_qp_minus26 out of range");
  ensure(p.pps_cb_qp_offset >= -12 && p.pps_cb_qp_offset <= 12 && p.pps_cr_qp_offset >= -12 && p.pps_cr_qp_offset <= 12,
         HEIC_E_BITSTREAM, "pps chroma qp offset out of range");
  return p;
}

static bool is_irap(uint32_t t) { return t >= 16 && t <= 23; }  // slice.rs:257-270
static bool is_idr(uint32_t t) { return t == 19 || t == 20; }

// slice.rs:44-204 (+ the byte offsets the seam needs).  Errors instead of the reference's
// assert!/unimplemented! (slice.rs:60-63,106-108).
heic_slice_header slice_segment_header(const uint8_t* rbsp, size_t len, uint32_t nal_unit_type,
                                       const heic_sps& sps, const heic_pps& pps,
                                       const uint32_t* epb_pos, size_t n_epb, bool raw_offsets) {
  RbspReader r(rbsp, len);
  heic_slice_header h;
  std::memset(&h, 0, sizeof h);
  h.first_slice_segment_in_pic_flag = r.read_flag();
  if (is_irap(nal_unit_type)) h.no_output_of_prior_pics_flag = r.read_flag();
  h.slice_pic_parameter_set_id = r.read_ue();
  ensure(h.first_slice_segment_in_pic_flag, HEIC_E_UNSUPPORTED,
         "first_slice_segment_in_pic_flag = 0: pictures with more than one slice segment are not supported");
  for (uint32_t i = 0; i < pps.num_extra_slice_header_bits; ++i) r.read_flag();
  h.slice_type = r.read_ue();
  ensure(h.slice_type <= 2, HEIC_E_BITSTREAM, "invalid slice_type");
  if (pps.output_flag_present_flag) r.read_flag();
  if (sps.separate_colour_plane_flag) r.read_u8(2);
  ensure(is_idr(nal_unit_type), HEIC_E_UNSUPPORTED, "only IDR pictures are supported (HEIC still images)");
  if (sps.sample_adaptive_offset_enabled_flag) {
    h.slice_sao_luma_flag = r.read_flag();
    uint32_t chroma_array_type = sps.separate_colour_plane_flag ? 0 : sps.chroma_format_idc;
    if (chroma_array_type != 0) h.slice_sao_chroma_flag = r.read_flag();
  }
  ensure(h.slice_type == 2, HEIC_E_UNSUPPORTED, "P/B slices are not supported (intra still images only)");
  h.slice_qp_delta = r.read_se();
  if (pps.pps_slice_chroma_qp_offsets_present_flag) {
    h.slice_cb_qp_offset = r.read_se();
    h.slice_cr_qp_offset = r.read_se();
  }
  if (pps.deblocking_filter_override_enabled_flag) h.deblocking_filter_override_flag = r.read_flag();
  h.slice_deblocking_filter_disabled_flag = pps.pps_deblocking_filter_disabled_flag;
  h.slice_beta_offset_div2 = pps.pps_beta_offset_div2;
  h.slice_tc_offset_div2 = pps.pps_tc_offset_div2;
  if (h.deblocking_filter_override_flag) {
    h.slice_deblocking_filter_disabled_flag = r.read_flag();
    if (!h.slice_deblocking_filter_disabled_flag) {
      h.slice_beta_offset_div2 = r.read_se();
      h.slice_tc_offset_div2 = r.read_se();
    }
  }
  h.slice_loop_filter_across_slices_enabled_flag = pps.pps_loop_filter_across_slices_enabled_flag;
  if (pps.pps_loop_filter_across_slices_enabled_flag &&
      (h.slice_sao_luma_flag || h.slice_sao_chroma_flag || !h.slice_deblocking_filter_disabled_flag))
    h.slice_loop_filter_across_slices_enabled_flag = r.read_flag();
  if (pps.tiles_enabled_flag || pps.entropy_coding_sync_enabled_flag) {
    h.num_entry_point_offsets = r.read_ue();
    ensure(h.num_entry_point_offsets <= HEIC_MAX_ENTRY_POINTS, HEIC_E_UNSUPPORTED, "too many entry points");
    if (h.num_entry_point_offsets > 0) {
      uint32_t offset_len_minus1 = r.read_ue();
      ensure(offset_len_minus1 < 32, HEIC_E_BITSTREAM, "offset_len_minus1 > 31");
      for (uint32_t i = 0; i < h.num_entry_point_offsets; ++i)
        h.entry_point_offset_minus1[i] = r.read_u32(offset_len_minus1 + 1);
    }
  }
  if (pps.slice_segment_header_extension_present_flag) {
    uint32_t n = r.read_ue();
    for (uint32_t i = 0; i < n; ++i) r.read_u8(8);
  }
  r.byte_alignment();
  int qp = 26 + pps.init_qp_minus26 + h.slice_qp_delta;
  ensure(qp >= 0 && qp <= 51, HEIC_E_BITSTREAM, "SliceQpY out of range");

  h.slice_data_byte_offset = static_cast<uint32_t>(r.byte_position());
  if (raw_offsets) {
    // keep everything in raw (escaped) byte counts: the GPU removes the emulation prevention bytes and re-bases
    uint64_t e = h.slice_data_byte_offset;
    for (size_t i = 0; i < n_epb; ++i) {
      if (epb_pos[i] <= e) ++e;
      else break;
    }
    h.slice_data_byte_offset = static_cast<uint32_t>(e);
    h.substream_offset[0] = 0;
    uint64_t sum = 0;
    for (uint32_t k = 0; k < h.num_entry_point_offsets; ++k) {
      sum += uint64_t{h.entry_point_offset_minus1[k]} + 1;
      if (sum > 0xffffffffull) bail(HEIC_E_BITSTREAM, "entry point beyond slice data");
      h.substream_offset[k + 1] = static_cast<uint32_t>(sum);
    }
    return h;
  }
  // Entry points count bytes of the ESCAPED NAL (7.4.7.1).  Map: un-escaped u -> escaped e, then
  // boundaries back to un-escaped positions.  epb_pos are positions in the escaped payload.
  uint64_t e = h.slice_data_byte_offset;
  for (size_t i = 0; i < n_epb; ++i) {
    if (epb_pos[i] <= e) ++e;
    else break;
  }
  h.substream_offset[0] = 0;
  uint64_t boundary = e;
  for (uint32_t k = 0; k < h.num_entry_point_offsets; ++k) {
    boundary += uint64_t{h.entry_point_offset_minus1[k]} + 1;
    size_t removed = std::lower_bound(epb_pos, epb_pos + n_epb, static_cast<uint32_t>(std::min<uint64_t>(boundary, 0xffffffffu))) - epb_pos;
    uint64_t u = boundary - removed;
    if (u < h.slice_data_byte_offset || u > len) bail(HEIC_E_BITSTREAM, "entry point beyond slice data");
    h.substream_offset[k + 1] = static_cast<uint32_t>(u - h.slice_data_byte_offset);
    if (h.substream_offset[k + 1] <= h.substream_offset[k] && k + 1 > 0 && h.substream_offset[k + 1] < h.substream_offset[k])
      bail(HEIC_E_BITSTREAM, "entry points not monotonic");
  }
  return h;
}

heic_slice_header slice_segment_header_raw(const uint8_t* nal_payload, size_t len, uint32_t nal_unit_type,
                                           const heic_sps& sps, const heic_pps& pps) {
  // the header is a few dozen bytes (4 per entry point at most): un-escape a bounded prefix only
  const size_t prefix = std::min<size_t>(len, 64 + 5 * static_cast<size_t>(HEIC_MAX_ENTRY_POINTS));
  std::vector<uint32_t> epb;
  std::vector<uint8_t> head = RbspReader::remove_emulation_prevention(nal_payload, prefix, &epb);
  // a 00 00 03 cut by the prefix end is not a removal candidate the full scan would agree on; the header ends well before
  return slice_segment_header(head.data(), head.size(), nal_unit_type, sps, pps, epb.data(), epb.size(), true);
}

}  // namespace heic
