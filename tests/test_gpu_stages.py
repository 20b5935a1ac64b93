"""GPU parity, stage by stage, through the C ABI: every intermediate buffer of the CUDA pipeline against the
CPU oracle (oracle/hevc_oracle.c, itself pinned to FFmpeg) on all 48 tiles of halfmoonbay.heic.  Bit-exact."""
import numpy as np
import pytest

import heif_b200 as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def batch(decoder, heic_file):
    b = decoder.batch([heic_file.primary])
    yield b
    b.close()


def _check_status(batch, oracle_tiles):
    st = batch.status()
    for t in range(batch.n_tiles):
        assert st[t].code == 0, f"tile {t}: status {st[t].code}"
        ref = oracle_tiles(t)
        assert st[t].bins_decoded == ref["bins"], f"tile {t}: bins"
        assert st[t].ctus_decoded == ref["ctus"]


def test_stage_by_stage(batch, oracle_tiles):
    n = batch.n_tiles
    # 1. CABAC: syntax + TransCoeffLevel
    batch.run(H.STAGE_CABAC)
    batch.sync()
    _check_status(batch, oracle_tiles)
    for t in range(n):
        d, ref = batch.dump_tile(t), oracle_tiles(t)
        assert np.array_equal(d["tu_map"], ref["tu_map"]), f"tile {t}: tu_map"
        for c in range(3):
            assert np.array_equal(d["coeff"][c], ref["level"][c]), f"tile {t}: levels c={c}"
        assert np.array_equal(d["qp_map"], ref["qp_map"]), f"tile {t}: qp_map"
        assert np.array_equal(d["sao"], ref["sao"]), f"tile {t}: sao"
    # 2. scaling + inverse transform -> residual (compared where a block is coded)
    batch.run(H.STAGE_TRANSFORM)
    batch.sync()
    for t in range(n):
        d, ref = batch.dump_tile(t), oracle_tiles(t)
        for c in range(3):
            mask = ref["resid"][c] != 0
            assert np.array_equal(d["coeff"][c][mask], ref["resid"][c][mask]), f"tile {t}: residual c={c}"
            diff = np.flatnonzero(d["coeff"][c] != ref["resid"][c])
            assert diff.size == 0, f"tile {t}: residual c={c} differs at {diff[:8]}"
    # 3. intra prediction + reconstruction
    batch.run(H.STAGE_INTRA)
    batch.sync()
    for t in range(n):
        d, ref = batch.dump_tile(t), oracle_tiles(t)
        for c in range(3):
            assert np.array_equal(d["plane"][c], ref["recon"][c]), f"tile {t}: recon c={c}"
    # 4. deblocking
    batch.run(H.STAGE_DEBLOCK)
    batch.sync()
    for t in range(n):
        d, ref = batch.dump_tile(t), oracle_tiles(t)
        for c in range(3):
            assert np.array_equal(d["plane"][c], ref["deblocked"][c]), f"tile {t}: deblocked c={c}"
    # 5. SAO
    batch.run(H.STAGE_SAO)
    batch.sync()
    for t in range(n):
        d, ref = batch.dump_tile(t), oracle_tiles(t)
        for c in range(3):
            assert np.array_equal(d["plane"][c], ref["plane"][c]), f"tile {t}: final c={c}"
