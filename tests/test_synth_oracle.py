"""Synthetic streams (tests/synth: the deterministic in-repo CABAC encoder) through the CPU oracle, against the committed
FFmpeg golden hashes: pins the oracle on every feature the fixture does not exercise (CTB 16/64, no WPP, transform skip,
sign data hiding, transform hierarchy, explicit scaling lists, 4:0:0, ragged picture sizes, ...)."""
import hashlib
import json
import os

import pytest

from oracle import oracle_py as O
from tests.synth import synth
from tests.synth.configs import CONFIGS, SEEDS

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "synth_hashes.json")))["streams"]


@pytest.mark.parametrize("name,cfg", CONFIGS, ids=[n for n, _ in CONFIGS])
def test_oracle_matches_ffmpeg_on_synthetic_streams(built, name, cfg):
    for seed in SEEDS:
        pic = synth.encode(seed, **cfg)
        g = GOLDEN[f"{name}/{seed}"]
        assert hashlib.sha256(pic.annexb()).hexdigest() == g["annexb_sha256"], "the generator is not deterministic"
        t = pic.tile
        ref = O.decode_picture(pic.sps, pic.pps, t.header, (t.rbsp, t.rbsp_len), intermediates=False)
        assert [hashlib.sha256(p.tobytes()).hexdigest() for p in ref["plane"]] == g["planes"], (name, seed)


def test_generator_is_seed_sensitive(built):
    a, b = synth.encode(0), synth.encode(1)
    assert a.annexb() != b.annexb()
    assert synth.encode(0).annexb() == a.annexb()
