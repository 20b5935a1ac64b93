"""Pins the CPU oracle (oracle/hevc_oracle.c) to FFmpeg's independent HEVC decoder: per-tile plane SHA-256 of all 48
tiles of halfmoonbay.heic (tests/golden/fixture_hashes.json, produced by tests/golden/make_golden.py)."""
import hashlib
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from oracle import oracle_py as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "fixture_hashes.json")))


def test_oracle_matches_ffmpeg_on_every_fixture_tile(heic_file):
    img = heic_file.primary

    def one(t):
        td = img.tiles[t]
        r = O.decode_picture(img.sps, img.pps, td.header, (td.rbsp, td.rbsp_len), intermediates=False)
        return [hashlib.sha256(p.tobytes()).hexdigest() for p in r["plane"]], r["plane"]

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        res = list(ex.map(one, range(img.n_tiles)))
    for t, (hashes, _) in enumerate(res):
        assert hashes == GOLDEN["tiles"][t], f"tile {t}"
    h = hashlib.sha256()
    for _, planes in res:
        for p in planes:
            h.update(p.tobytes())
    assert h.hexdigest() == GOLDEN["all_tiles_concat"]
    # stitched + cropped planes
    canvas = [np.zeros((3072, 4096), np.uint8), np.zeros((1536, 2048), np.uint8), np.zeros((1536, 2048), np.uint8)]
    for t, (_, planes) in enumerate(res):
        r, c = divmod(t, 8)
        for i, p in enumerate(planes):
            s = 512 >> (1 if i else 0)
            canvas[i][r * s:(r + 1) * s, c * s:(c + 1) * s] = p
    st = [canvas[0][:3024, :4032], canvas[1][:1512, :2016], canvas[2][:1512, :2016]]
    assert [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in st] == GOLDEN["stitched"]


def test_context_init_known_answers(built):
    """9.3.2.2 (arithmetic.rs:40-78): initValue 154 is the 'equiprobable' value: preCtxState 64 at any QP -> state 0, MPS 1."""
    for qp in (0, 15, 26, 51):
        st = O.context_init(qp)
        assert st[5] == 1 and st[18] == 1 and st[19] == 1  # cu_transquant_bypass, cu_qp_delta_abs x2 (initValue 154)
    st = O.context_init(15)
    # sao_merge (153): slope 9 -> m = 0, offset 9 -> n = 56: pre = 56 -> MPS 0, pState 7
    assert st[0] == (7 << 1)
    # split_cu_flag[0] (139): m = -5, n = 72: pre = ((-5 * 15) >> 4) + 72 = 67 -> MPS 1, pState 3
    assert st[2] == ((3 << 1) | 1)


def test_inverse_transform_known_answers(built):
    """8.6.4.2: a DC-only block reconstructs to a constant; the 4x4 DST of a DC-only block is not constant."""
    for log2 in (2, 3, 4, 5):
        n = 1 << log2
        c = np.zeros((n, n), np.int16)
        c[0, 0] = 64
        r = O.idct(c.ravel(), log2).reshape(n, n)
        # (64*64 + 64) >> 7 = 32 -> (32*64 + 2048) >> 12 = 1
        assert (r == 1).all(), log2
    c = np.zeros(16, np.int16)
    c[0] = 64
    r = O.idct(c, 2, dst=True)
    assert len(set(r.tolist())) > 1
    # the two passes commute with transposition for the symmetric DCT basis
    rng = np.random.default_rng(0)
    a = rng.integers(-256, 256, (8, 8)).astype(np.int16)
    assert np.array_equal(O.idct(a.ravel(), 3).reshape(8, 8).T, O.idct(np.ascontiguousarray(a.T).ravel(), 3).reshape(8, 8))


def test_colour_definition_corner_cases(built):
    """the frozen integer definition at the corners of the YCbCr cube (full range BT.601)."""
    def px(y, cb, cr):
        planes = np.concatenate([np.full(64, y, np.uint8), np.full(16, cb, np.uint8), np.full(16, cr, np.uint8)])
        return tuple(int(v) for v in O.color_stitch(planes, 1, 1, 8, 8, 8, 8)[0, 0])

    assert px(0, 128, 128) == (0, 0, 0)
    assert px(255, 128, 128) == (255, 255, 255)
    assert px(128, 128, 128) == (128, 128, 128)
    assert px(128, 128, 255) == (255, 128 + ((-183 * 127 + 128) >> 8), 128)
    assert px(128, 255, 128) == (128, 128 + ((-88 * 127 + 128) >> 8), 255)
    assert px(255, 0, 0) == (255 + ((359 * -128 + 128) >> 8), 255, 255 + ((454 * -128 + 128) >> 8))
    assert px(0, 0, 0)[0] == 0 and px(0, 0, 0)[2] == 0  # clipped
