"""Tile pool of SURVEY section 8(d) config 5: the fixture's 48 real 512x512 tiles + N synthetic 512x512 WPP tiles that decode
under the fixture's own SPS/PPS, and the seeded composition of 8x6 grid images from it.  TEST / BENCH INFRASTRUCTURE ONLY
(the synthetic tiles come from tests/synth/hevc_synth.cc, the in-repo CABAC encoder)."""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import heif_b200 as H
from heif_b200 import _capi as K

from . import synth

# The fixture's coding configuration: CTB 32, min CB 8, TB 4..32, default scaling lists, SAO, WPP,
# cu_qp_delta with depth 2, init_qp -11, chroma QP offsets +2.  A synthetic tile generated with it parses and decodes
# with the fixture's SPS/PPS, so real and synthetic tiles can share one grid image.
FIXTURE_CODING = dict(width=512, height=512, chroma_format_idc=1, log2_min_cb=3, log2_ctb=5, log2_min_tb=2, log2_max_tb=5,
                      max_transform_hierarchy_depth_intra=0, scaling_list_mode=1, sao=1, cu_qp_delta=1, diff_cu_qp_delta_depth=2,
                      init_qp_minus26=-11, cb_qp_offset=2, cr_qp_offset=2, wpp=1)


def synth_params(k: int) -> dict:
    """Generator knobs of synthetic tile k.  A sweep sized to the fixture's bimodal 1.4-76 KB spread: 3 of 8 tiles are light
    (a few KB and mostly unsplit coding units, like the fixture's sky tiles), 5 of 8 heavy (20-90 KB); every tile has its
    own seed.  SAO is switched on for about 3 % of the CTBs, as in the fixture (0-4 % per tile); left to the random walk it
    would be on for most CTBs, which no encoder produces."""
    if k % 8 < 3:
        return dict(slice_qp_delta=(20, 24, 28)[k % 3], lps_gain=(0.4, 0.5, 0.6, 0.7, 0.8)[(k // 8) % 5],
                    split_cu_prob=(0.15, 0.25, 0.4)[(k // 3) % 3], sao_on_prob=0.03)
    return dict(slice_qp_delta=(0, 2, 4, 6, 8, 10)[(k // 8) % 6], lps_gain=(0.8, 0.9, 1.0, 1.1, 1.2)[k % 5], sao_on_prob=0.03)


class Pool:
    def __init__(self):
        self.descs = []     # K.TileDesc per pool entry (pointing into self.keep buffers)
        self.kind = []      # "real" / "synth"
        self.keep = []
        self.sps = None
        self.pps = None

    def __len__(self):
        return len(self.descs)

    def sizes(self):
        return np.array([d.rbsp_len for d in self.descs], np.int64)

    def composition(self) -> dict:
        sz = self.sizes()
        real = np.array([k == "real" for k in self.kind])
        q = lambda a: [int(x) for x in np.percentile(a, [0, 25, 50, 75, 100])]
        return {"real_tiles": int(real.sum()), "synthetic_tiles": int((~real).sum()),
                "slice_bytes_quartiles_real": q(sz[real]), "slice_bytes_quartiles_synthetic": q(sz[~real]) if (~real).any() else None}


def build_pool(heic_file, n_synth: int = 512, threads: int | None = None) -> Pool:
    base = heic_file.primary
    pool = Pool()
    pool.sps, pool.pps = base.sps, base.pps
    for t in range(base.n_tiles):
        pool.descs.append(base.tiles[t])
        pool.kind.append("real")
    pool.keep.append(heic_file)

    def one(k):
        return synth.encode_nals(1000 + k, **synth_params(k), **FIXTURE_CODING)

    synth._load()
    with ThreadPoolExecutor(threads or min(32, os.cpu_count() or 1)) as ex:
        nals = list(ex.map(one, range(n_synth)))
    for vps, sps_nal, pps_nal, slice_nal in nals:
        rbsp, epb = H.remove_emulation_prevention(slice_nal[2:], with_positions=True)
        buf = (K.u8 * len(rbsp)).from_buffer_copy(rbsp)
        td = K.TileDesc()
        td.rbsp = C.cast(buf, C.POINTER(K.u8))
        td.rbsp_len = len(rbsp)
        td.nal_unit_type = 20
        td.header = H.parse_slice_header(rbsp, 20, base.sps, base.pps, epb)  # parsed under the FIXTURE's parameter sets
        pool.descs.append(td)
        pool.kind.append("synth")
        pool.keep.append(buf)
    return pool


def image_tile_ids(pool_size: int, image_idx: int, n_tiles: int = 48, seed: int = 1) -> np.ndarray:
    """Pool entries of global image `image_idx`: n_tiles distinct entries drawn by rng(seed, image_idx), so that any rank
    can compose any image of the job without generating the others."""
    rng = np.random.default_rng([seed, image_idx])
    return rng.choice(pool_size, size=n_tiles, replace=False)


def compose_images(pool: Pool, base_image, image_indices, seed: int = 1):
    """-> (list of K.ImageDesc, keep-alive list, ids array [n_images, n_tiles] of pool entries)."""
    n_tiles = base_image.n_tiles
    images, keep, ids = [], [], []
    tsize = C.sizeof(K.TileDesc)
    for gi in image_indices:
        sel = image_tile_ids(len(pool), int(gi), n_tiles, seed)
        tiles = (K.TileDesc * n_tiles)()
        for d, s in enumerate(sel):
            C.memmove(C.byref(tiles, d * tsize), C.byref(pool.descs[int(s)]), tsize)
        im = K.ImageDesc()
        C.memmove(C.byref(im), C.byref(base_image), C.sizeof(K.ImageDesc))
        im.tiles = C.cast(tiles, C.POINTER(K.TileDesc))
        keep.append(tiles)
        images.append(im)
        ids.append(sel)
    return images, keep, np.array(ids, np.int64)
