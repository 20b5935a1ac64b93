"""The reference's own known-answer tests, restated against this repo's host layer and the CPU oracle.

    src/hevc/rbsp_reader.rs:144-184      ue(v)/se(v) Tables 9-2 / 9-3
    src/hevc/rbsp_reader.rs:186-303      12 emulation-prevention cases
    src/cabac/decoder.rs:309-373         TR (Table 9-39) and intra_chroma_pred_mode (Table 9-41) bin strings
    tests/libheif_comparison.rs:102-111  container metadata of halfmoonbay.heic (values libheif reports)
"""
import pytest

import heif_b200 as H
from oracle import oracle_py as O


@pytest.mark.parametrize("byte,expected", [(0b10000000, 0), (0b01000000, 1), (0b01100000, 2), (0b00100000, 3), (0b00101000, 4),
                                           (0b00110000, 5), (0b00111000, 6), (0b00010000, 7), (0b00010010, 8), (0b00010100, 9)])
def test_ue_table_9_2(built, byte, expected):
    assert H.read_ue(bytes([byte]))[0] == expected


@pytest.mark.parametrize("byte,expected", [(0b10000000, 0), (0b01000000, 1), (0b01100000, -1), (0b00100000, 2), (0b00101000, -2),
                                           (0b00110000, 3), (0b00111000, -3)])
def test_se_table_9_3(built, byte, expected):
    assert H.read_se(bytes([byte]))[0] == expected


EPB_CASES = [
    ([0x01, 0x02, 0x03, 0x04, 0x05], None),                                      # no pattern
    ([0x01, 0x00, 0x02], None),                                                  # single zero
    ([0x01, 0x00, 0x00, 0x04], None),                                            # double zero
    ([0x00, 0x00, 0x03, 0x00], [0x00, 0x00, 0x00]),                              # basic
    ([0x00, 0x00, 0x03, 0x01], [0x00, 0x00, 0x01]),
    ([0x00, 0x00, 0x03, 0x02], [0x00, 0x00, 0x02]),
    ([0x00, 0x00, 0x03, 0x03], [0x00, 0x00, 0x03]),
    ([0x00, 0x00, 0x03, 0x04], None),                                            # 03 kept: the reference's pinned quirk
    ([0x01, 0x00, 0x00, 0x03], [0x01, 0x00, 0x00]),                              # at end
    ([0x00, 0x00, 0x03, 0x00, 0xFF, 0x00, 0x00, 0x03, 0x01], [0x00, 0x00, 0x00, 0xFF, 0x00, 0x00, 0x01]),
    ([0x42, 0x01, 0x01, 0x03, 0x70, 0x00, 0x00, 0x03, 0x00], [0x42, 0x01, 0x01, 0x03, 0x70, 0x00, 0x00, 0x00]),
    ([0x01, 0x03, 0x70, 0x00, 0x00, 0x03, 0x00, 0xb0, 0x00, 0x00, 0x03, 0x00, 0x00, 0x03, 0x00, 0x5a, 0xa0, 0x04],
     [0x01, 0x03, 0x70, 0x00, 0x00, 0x00, 0xb0, 0x00, 0x00, 0x00, 0x00, 0x00, 0x5a, 0xa0, 0x04]),   # real SPS bytes
    ([0x00, 0x00, 0x03, 0x00, 0x00, 0x03, 0x01], [0x00, 0x00, 0x00, 0x00, 0x01]),  # consecutive
    ([], None),                                                                  # empty
]


@pytest.mark.parametrize("data,expected", EPB_CASES)
def test_remove_emulation_prevention(built, data, expected):
    out = H.remove_emulation_prevention(bytes(data))
    assert list(out) == (data if expected is None else expected)


def test_emulation_prevention_positions(built):
    out, pos = H.remove_emulation_prevention(bytes([0, 0, 3, 0, 0xFF, 0, 0, 3, 1]), with_positions=True)
    assert pos == [2, 7] and len(out) == 7


def test_tr_table_9_39_unary(built):  # cMax = 5, cRiceParam = 0
    for value, bins in enumerate([[0], [1, 0], [1, 1, 0], [1, 1, 1, 0], [1, 1, 1, 1, 0], [1, 1, 1, 1, 1]]):
        assert O.test_binarization(0, 5, bins) == (value, len(bins))


def test_intra_chroma_pred_mode_table_9_41(built):
    for value, bins in [(4, [0]), (0, [1, 0, 0]), (1, [1, 0, 1]), (2, [1, 1, 0]), (3, [1, 1, 1])]:
        assert O.test_binarization(1, 0, bins) == (value, len(bins))


def test_egk_and_coeff_abs_level_remaining_known_answers(built):
    # 9.3.3.6 EGk: k = 0: "0" -> 0, "100" -> 1, "101" -> 2, "11000" -> 3; k = 1: "00" -> 0, "01" -> 1, "1000" -> 2
    assert O.test_binarization(2, 0, [0])[0] == 0
    assert O.test_binarization(2, 0, [1, 0, 0])[0] == 1
    assert O.test_binarization(2, 0, [1, 0, 1])[0] == 2
    assert O.test_binarization(2, 0, [1, 1, 0, 0, 0])[0] == 3
    assert O.test_binarization(2, 1, [0, 1])[0] == 1
    assert O.test_binarization(2, 1, [1, 0, 0, 0])[0] == 2
    # 9.3.3.11 coeff_abs_level_remaining, rice 0: prefix < 4 is the value; "11110" -> 4 (EG1 escape starts at 4)
    for v in range(4):
        assert O.test_binarization(3, 0, [1] * v + [0])[0] == v
    assert O.test_binarization(3, 0, [1, 1, 1, 1, 0, 0])[0] == 4
    assert O.test_binarization(3, 0, [1, 1, 1, 1, 0, 1])[0] == 5
    assert O.test_binarization(3, 0, [1, 1, 1, 1, 1, 0, 0, 0])[0] == 6
    # rice 2: "10" + FL(2) "11" -> (1 << 2) + 3 = 7
    assert O.test_binarization(3, 2, [1, 0, 1, 1])[0] == 7


def test_container_metadata_like_libheif_comparison(heic_file):
    """tests/libheif_comparison.rs:102-111 — expected values are what libheif reports for the fixture."""
    info = heic_file.info
    assert (info.ispe_width, info.ispe_height) == (4032, 3024)
    assert info.rotation_ccw_quarter_turns == 3
    assert (info.rotated_width, info.rotated_height) == (3024, 4032)
    assert (info.luma_bits, info.chroma_bits) == (8, 8)
    assert info.primary_item_id == 49 and info.is_grid == 1
    assert info.thumbnail_count == 0
    assert info.item_count == 53


def test_parameter_sets_of_the_fixture(heic_file):
    """the parameter-set facts the kernels' launch parameters come from."""
    img = heic_file.primary
    s, p = img.sps, img.pps
    assert (s.pic_width_in_luma_samples, s.pic_height_in_luma_samples) == (512, 512)
    assert s.chroma_format_idc == 1 and s.bit_depth_luma_minus8 == 0
    assert s.log2_min_luma_coding_block_size_minus3 == 0 and s.log2_diff_max_min_luma_coding_block_size == 2   # CTB 32
    assert s.log2_min_luma_transform_block_size_minus2 == 0 and s.log2_diff_max_min_luma_transform_block_size == 3
    assert s.max_transform_hierarchy_depth_intra == 0
    assert s.scaling_list_enabled_flag == 1 and s.sps_scaling_list_data_present_flag == 0
    assert s.sample_adaptive_offset_enabled_flag == 1 and s.pcm_enabled_flag == 0 and s.amp_enabled_flag == 0
    assert s.strong_intra_smoothing_enabled_flag == 0
    assert s.video_full_range_flag == 1 and s.matrix_coeffs == 6
    assert p.init_qp_minus26 == -11 and p.cu_qp_delta_enabled_flag == 1 and p.diff_cu_qp_delta_depth == 2
    assert (p.pps_cb_qp_offset, p.pps_cr_qp_offset) == (2, 2)
    assert p.entropy_coding_sync_enabled_flag == 1 and p.tiles_enabled_flag == 0
    assert p.sign_data_hiding_enabled_flag == 0 and p.transform_skip_enabled_flag == 0
    assert (img.grid_rows, img.grid_cols, img.output_width, img.output_height) == (6, 8, 4032, 3024)
    assert img.n_tiles == 48 and img.rotation_ccw_quarter_turns == 3


def test_slice_headers_and_substreams_of_the_fixture(heic_file):
    """All 48 slices: IDR_N_LP I-slices with 15 entry points; every substream ends in a stop bit (SURVEY Appendix B #2:
    only correctly EPB-adjusted boundaries end in '1' + zero padding)."""
    img = heic_file.primary
    total = 0
    for t in range(img.n_tiles):
        td = img.tiles[t]
        h = td.header
        assert td.nal_unit_type == 20 and h.slice_type == 2 and h.first_slice_segment_in_pic_flag == 1
        assert h.slice_sao_luma_flag == 1 and h.slice_sao_chroma_flag == 1 and h.slice_qp_delta == 0
        assert h.num_entry_point_offsets == 15 and h.substream_offset[0] == 0
        rbsp = bytes(td.rbsp[: td.rbsp_len])
        base = h.slice_data_byte_offset
        ends = [base + h.substream_offset[k] for k in range(1, 16)] + [len(rbsp)]
        for e in ends:
            assert rbsp[e - 1] != 0, f"tile {t}: substream ending at {e} has no stop bit"
        total += td.rbsp_len
    assert 1_690_000 < total < 1_710_000  # 1.70 MB of slice data (minus NAL headers / EPBs)


def test_unsupported_and_malformed_inputs_become_error_codes(built):
    """The reference panics (todo!/unimplemented!/assert!) where this boundary returns codes."""
    with pytest.raises(H.HeicError) as e:
        H.HeicFile(b"\x00\x00\x00\x08ftyp")
    assert e.value.code in (H._capi.HEIC_E_BITSTREAM, H._capi.HEIC_E_UNSUPPORTED, H._capi.HEIC_E_INVALID_ARG)
    with pytest.raises(H.HeicError):
        H.parse_sps(b"\x01")
    with pytest.raises(H.HeicError):
        H.read_ue(b"\x00\x00\x00\x00\x00")  # no terminating 1 bit


def test_raw_slice_header_offsets_rebase_to_the_unescaped_ones(heic_file):
    """heic_b200_parse_slice_header_raw keeps raw byte counts (7.4.7.1); removing the emulation prevention bytes before
    each boundary must give the un-escaped offsets the converted header carries (what unescape_kernel does on the GPU)."""
    import heif_b200 as H
    import bisect

    img, raw = heic_file.primary, heic_file.primary_raw
    assert raw.n_tiles == img.n_tiles == 48
    n_with_epb = 0
    for t in range(48):
        a, b = img.tiles[t], raw.tiles[t]
        assert b.escaped == 1 and a.escaped == 0
        payload = bytes(b.rbsp[: b.rbsp_len])
        rbsp, epb = H.remove_emulation_prevention(payload, with_positions=True)
        assert rbsp == bytes(a.rbsp[: a.rbsp_len])
        n_with_epb += bool(epb)
        removed = lambda x: bisect.bisect_left(epb, x)
        e = b.header.slice_data_byte_offset
        data_off = e - removed(e)
        assert data_off == a.header.slice_data_byte_offset
        n = a.header.num_entry_point_offsets
        assert b.header.num_entry_point_offsets == n
        for k in range(n + 1):
            boundary = e + b.header.substream_offset[k]
            assert boundary - removed(boundary) - data_off == a.header.substream_offset[k]
    assert n_with_epb == 2  # exactly two fixture tiles carry an emulation prevention byte
