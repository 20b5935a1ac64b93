import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixture_bytes():
    with open(os.path.join(ROOT, "tests", "golden", "halfmoonbay.heic"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def built():
    """Builds libheic_b200.so + the oracle if they are missing (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    if not os.path.exists(g.LIB):
        g.build()
    from oracle import oracle_py

    oracle_py.load()
    return g


@pytest.fixture(scope="session")
def heic_file(built, fixture_bytes):
    import heif_b200

    return heif_b200.HeicFile(fixture_bytes)


@pytest.fixture(scope="session")
def oracle_tiles(heic_file):
    """CPU-oracle decode (with per-stage intermediates) of the 48 fixture tiles, computed lazily."""
    from oracle import oracle_py

    cache = {}
    img = heic_file.primary

    def get(t):
        if t not in cache:
            td = img.tiles[t]
            cache[t] = oracle_py.decode_picture(img.sps, img.pps, td.header, (td.rbsp, td.rbsp_len))
        return cache[t]

    return get


@pytest.fixture(scope="session", params=["auto", "thread_per_substream"])
def decoder(built, request):
    """Both CABAC mappings: 'auto' picks warp-per-substream for the small loads the tests use; the second context forces
    the 32-tiles-per-CTA thread-per-substream kernel that large batches (bench.py) run."""
    import heif_b200

    knobs = {}
    if request.param == "thread_per_substream":
        # ... and with two persistent CTAs, so that every load of more than 64 tiles exercises the warp-by-warp
        # hand-over from one group of tiles to the next
        knobs = {"HEIC_B200_LOW_LATENCY_TILES": "0", "HEIC_B200_CABAC_RESIDENT": "2"}
    old = {k: os.environ.get(k) for k in knobs}
    os.environ.update(knobs)
    try:
        dec = heif_b200.HeicDecoder(device=0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    yield dec
    dec.close()
