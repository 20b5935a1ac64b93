"""The bench workload: tile pool + seeded image composition + image_idx % n_gpus sharding."""
import numpy as np
import pytest

from heif_b200 import sharding


@pytest.fixture(scope="module")
def small_pool(heic_file):
    from tests.synth import pool as P

    return P.build_pool(heic_file, 24)


def test_pool_is_deterministic_and_decodes_under_the_fixture_parameter_sets(small_pool, heic_file):
    from oracle import oracle_py as O
    from tests.synth import pool as P

    again = P.build_pool(heic_file, 24)
    assert [d.rbsp_len for d in again.descs] == [d.rbsp_len for d in small_pool.descs]
    assert small_pool.kind.count("real") == 48 and small_pool.kind.count("synth") == 24
    for k in (48, 49, 53, 71):  # light and heavy synthetic tiles, parsed and decoded with the FIXTURE's SPS/PPS
        td = small_pool.descs[k]
        assert bytes(td.rbsp[: td.rbsp_len]) == bytes(again.descs[k].rbsp[: td.rbsp_len])
        r = O.decode_picture(small_pool.sps, small_pool.pps, td.header, (td.rbsp, td.rbsp_len), intermediates=False)
        assert r["ctus"] == 256
    sz = small_pool.sizes()[48:]
    assert sz.min() < 6000 and sz.max() > 20000  # the sweep spans the fixture's light and heavy tiles


def test_image_composition_is_per_image_seeded_and_duplicate_free():
    from tests.synth import pool as P

    a = P.image_tile_ids(560, 7)
    assert np.array_equal(a, P.image_tile_ids(560, 7)) and not np.array_equal(a, P.image_tile_ids(560, 8))
    assert len(set(a.tolist())) == 48 and a.max() < 560
    # rank r of n composes exactly the images i with i % n == r, whatever the other ranks do
    assert list(sharding.shard_modulo(10, 1, 4)) == [1, 5, 9]
    assert sorted(sum((list(sharding.shard_modulo(11, r, 3)) for r in range(3)), [])) == list(range(11))
    with pytest.raises(ValueError):
        sharding.shard_modulo(4, 4, 4)


@pytest.mark.gpu
def test_bench_scale_batch_matches_oracle_pixels(built, heic_file):
    """>= 2048 tiles through the throughput path with its default persistent CTAs (no RESIDENT override, the R = 1 intra
    path, size-dealt CABAC groups): random images against the oracle, pixel by pixel; no warp holds two copies of a tile."""
    import bench
    import heif_b200 as H
    from tests.synth import pool as P

    pool = P.build_pool(heic_file, 128)
    n_img = 44  # 2112 tiles
    images, keep, ids = P.compose_images(pool, heic_file.primary, range(n_img), 1)
    with H.HeicDecoder(device=0) as dec:
        batch = dec.batch(images)
        info = bench.check_groups_distinct(batch, ids.reshape(-1))
        assert info["tiles_per_group"] == 32 and info["groups_with_duplicates"] == 0
        batch.decode()
        batch.sync()
        st = batch.status()
        assert all(st[i].code == 0 for i in range(batch.n_tiles))
        chk = np.array([0, 13, 29, 43])
        planes = np.zeros((len(chk), 48, bench.TILE_BYTES), np.uint8)
        rgb = np.zeros((len(chk), bench.OUT_H, bench.OUT_W, 3), np.uint8)
        bench.CpuDecoder(pool, heic_file.primary, 8).decode(ids[chk], planes, rgb)
        for k, i in enumerate(chk):
            assert np.array_equal(batch.download_image(int(i)), rgb[k]), f"image {i}"
        batch.close()
