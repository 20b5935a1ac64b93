"""GPU parity on the synthetic streams: every configuration of tests/synth/configs.py, stage by stage against the CPU
oracle and end to end against the committed FFmpeg hashes; then all of them in ONE heterogeneous batch."""
import hashlib
import json
import os

import numpy as np
import pytest

import heif_b200 as H
from oracle import oracle_py as O
from tests.synth import synth
from tests.synth.configs import CONFIGS, SEEDS

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "synth_hashes.json")))["streams"]


@pytest.mark.parametrize("name,cfg", CONFIGS, ids=[n for n, _ in CONFIGS])
def test_stages_match_oracle(decoder, name, cfg):
    pic = synth.encode(SEEDS[0], **cfg)
    t = pic.tile
    ref = O.decode_picture(pic.sps, pic.pps, t.header, (t.rbsp, t.rbsp_len))
    n_comp = 3 if pic.sps.chroma_format_idc else 1
    with decoder.batch([pic.desc]) as b:
        b.run(H.STAGE_CABAC)
        b.sync()
        st = b.status()
        assert st[0].code == 0 and st[0].bins_decoded == ref["bins"] and st[0].ctus_decoded == ref["ctus"]
        d = b.dump_tile(0)
        assert np.array_equal(d["tu_map"], ref["tu_map"])
        for c in range(n_comp):
            assert np.array_equal(d["coeff"][c], ref["level"][c]), f"levels {c}"
        assert np.array_equal(d["qp_map"], ref["qp_map"]) and np.array_equal(d["sao"], ref["sao"])
        b.run(H.STAGE_TRANSFORM)
        d = b.dump_tile(0)
        for c in range(n_comp):
            assert np.array_equal(d["coeff"][c], ref["resid"][c]), f"residual {c}"
        for stage, key in ((H.STAGE_INTRA, "recon"), (H.STAGE_DEBLOCK, "deblocked"), (H.STAGE_SAO, "plane")):
            b.run(stage)
            d = b.dump_tile(0)
            for c in range(n_comp):
                assert np.array_equal(d["plane"][c], ref[key][c]), f"{key} {c}"
        b.run(H.STAGE_COLOR)
        rgb = b.download_rgb()[0]
    w, h = pic.sps.pic_width_in_luma_samples, pic.sps.pic_height_in_luma_samples
    if n_comp == 3:
        planes = np.concatenate([p.ravel() for p in ref["plane"]])
        exp = O.color_stitch(planes, 1, 1, w, h, w, h, pic.sps.video_full_range_flag, pic.sps.matrix_coeffs)
        assert np.array_equal(rgb, exp)


def test_heterogeneous_batch_matches_ffmpeg_golden(decoder):
    """All configurations and seeds in one call: different CTB sizes, WPP on/off, 4:0:0 and 4:2:0 in the same launches."""
    pics, keys = [], []
    for name, cfg in CONFIGS:
        for seed in SEEDS:
            pics.append(synth.encode(seed, **cfg))
            keys.append(f"{name}/{seed}")
    res = decoder.decode_grids_yuv([p.desc for p in pics])
    for key, pic, planes in zip(keys, pics, res):
        n_comp = 3 if pic.sps.chroma_format_idc else 1
        got = [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in planes[:n_comp]]
        assert got == GOLDEN[key]["planes"], key


# (seed, config) pairs whose slice data contains an emulation prevention byte — found by search: the arithmetic coder's
# output is close to random, so 00 00 0x turns up about once per 4 MB
EPB_SEEDS = (1249, 1914, 3113, 3754, 4324, 5903, 6936, 7057)
EPB_CONFIG = dict(init_qp_minus26=-20, lps_gain=2.0, width=512, height=256, wpp=1)


def test_heterogeneous_batch_from_raw_nal_payloads(decoder):
    """The same call with every picture shipped as a raw NAL payload (heic_tile_desc::escaped = 1): emulation
    prevention removal and entry-point re-basing on the GPU, for every geometry of the set plus eight WPP pictures that
    carry an emulation prevention byte inside their slice data (so the later entry points move)."""
    pics, keys = [], []
    for name, cfg in CONFIGS:
        for seed in SEEDS:
            pics.append(synth.encode(seed, **cfg))
            keys.append(f"{name}/{seed}")
    epb_pics = [synth.encode(seed, **EPB_CONFIG) for seed in EPB_SEEDS]
    assert all(p.n_epb >= 1 for p in epb_pics)
    res = decoder.decode_grids_yuv([p.desc_raw for p in pics + epb_pics])
    for key, pic, planes in zip(keys, pics, res):
        n_comp = 3 if pic.sps.chroma_format_idc else 1
        got = [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in planes[:n_comp]]
        assert got == GOLDEN[key]["planes"], key
    for pic, planes in zip(epb_pics, res[len(pics):]):
        ref = O.decode_picture(pic.sps, pic.pps, pic.header, (pic.tile.rbsp, pic.tile.rbsp_len))
        for c in range(3):
            assert np.array_equal(planes[c], ref["plane"][c])
