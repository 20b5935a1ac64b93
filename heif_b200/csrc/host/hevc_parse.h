// Host-side HEVC header parsing: function names follow the reference's
// src/hevc/parameter_set_reader.rs and src/hevc/slice.rs.
#pragma once
#include "rbsp_reader.h"

namespace heic {

void default_scaling_list(heic_scaling_list& sl);
heic_sps sequence_parameter_set_rbsp(const uint8_t* data, size_t len);
heic_pps picture_parameter_set_rbsp(const uint8_t* data, size_t len);
heic_slice_header slice_segment_header(const uint8_t* rbsp, size_t len, uint32_t nal_unit_type,
                                       const heic_sps& sps, const heic_pps& pps,
                                       const uint32_t* epb_pos, size_t n_epb, bool raw_offsets = false);
// the same header from a raw NAL payload, offsets left in raw byte counts (heic_tile_desc::escaped = 1)
heic_slice_header slice_segment_header_raw(const uint8_t* nal_payload, size_t len, uint32_t nal_unit_type,
                                           const heic_sps& sps, const heic_pps& pps);

}  // namespace heic
