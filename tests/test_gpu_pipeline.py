"""GPU parity of the whole path through the reference-facing C ABI (host descriptors in, host pixels out), against the
committed FFmpeg golden hashes and the CPU oracle.  Bit-exact: every stage is integer."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import heif_b200 as H
from heif_b200 import _capi as K
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "fixture_hashes.json")))


def permuted_image(base, perm):
    tiles = (K.TileDesc * base.n_tiles)()
    for d, s in enumerate(perm):
        C.memmove(C.byref(tiles, d * C.sizeof(K.TileDesc)), C.byref(base.tiles[int(s)]), C.sizeof(K.TileDesc))
    im = K.ImageDesc()
    C.memmove(C.byref(im), C.byref(base), C.sizeof(K.ImageDesc))
    im.tiles = C.cast(tiles, C.POINTER(K.TileDesc))
    return im, tiles


@pytest.fixture(scope="module")
def oracle_rgb(heic_file, oracle_tiles):
    img = heic_file.primary
    planes = np.concatenate([np.concatenate([p.ravel() for p in oracle_tiles(t)["plane"]]) for t in range(img.n_tiles)])
    return O.color_stitch(planes, img.grid_rows, img.grid_cols, 512, 512, img.output_width, img.output_height,
                          img.sps.video_full_range_flag, img.sps.matrix_coeffs)


def test_yuv_matches_ffmpeg_golden(decoder, heic_file):
    (y, cb, cr), = decoder.decode_grids_yuv([heic_file.primary])
    assert y.shape == (3024, 4032) and cb.shape == (1512, 2016)
    got = [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in (y, cb, cr)]
    assert got == GOLDEN["stitched"]


def test_decode_file_rgb_matches_oracle(decoder, fixture_bytes, oracle_rgb):
    rgb = decoder.decode(fixture_bytes)
    assert rgb.shape == (3024, 4032, 3)
    assert np.array_equal(rgb, oracle_rgb)


@pytest.mark.parametrize("turns", [1, 2, 3])
def test_irot_rotation(decoder, heic_file, oracle_rgb, turns):
    img = K.ImageDesc()
    C.memmove(C.byref(img), C.byref(heic_file.primary), C.sizeof(K.ImageDesc))
    img.rotation_ccw_quarter_turns = turns
    rgb = decoder.decode_grids([img], apply_transforms=True)[0]
    assert np.array_equal(rgb, np.rot90(oracle_rgb, turns))


def test_file_rotation_is_the_irot_of_the_fixture(decoder, fixture_bytes, oracle_rgb):
    rgb = decoder.decode(fixture_bytes, apply_transforms=True)
    assert rgb.shape == (4032, 3024, 3)  # irot = 3 -> 3024 x 4032 (libheif_comparison.rs:69-74)
    assert np.array_equal(rgb, np.rot90(oracle_rgb, 3))


def test_batch_of_permuted_images_is_the_permutation_of_the_outputs(decoder, heic_file, oracle_rgb):
    """Size-independent property: tiles are independent pictures, so permuting tiles permutes 512x512 blocks of the output."""
    base = heic_file.primary
    rng = np.random.default_rng(7)
    perms = [rng.permutation(48) for _ in range(5)]
    imgs, keep = zip(*[permuted_image(base, p) for p in perms])
    out = decoder.decode_grids(list(imgs))
    ref_full = np.zeros((3072, 4096, 3), np.uint8)
    planes = None
    for k, perm in enumerate(perms):
        for d, s in enumerate(perm):
            r, c = divmod(d, 8)
            rs, cs = divmod(int(s), 8)
            y0, x0 = r * 512, c * 512
            h, w = min(512, 3024 - y0), min(512, 4032 - x0)
            hs, ws = min(512, 3024 - rs * 512), min(512, 4032 - cs * 512)
            hh, ww = min(h, hs), min(w, ws)
            assert np.array_equal(out[k, y0:y0 + hh, x0:x0 + ww], oracle_rgb[rs * 512:rs * 512 + hh, cs * 512:cs * 512 + ww]), (k, d)


def test_corrupt_tile_is_reported_per_tile_and_does_not_poison_the_batch(decoder, heic_file, oracle_rgb):
    base = heic_file.primary
    img, tiles = permuted_image(base, list(range(48)))
    bad = bytearray(bytes(base.tiles[10].rbsp[: base.tiles[10].rbsp_len]))
    off = base.tiles[10].header.slice_data_byte_offset
    for i in range(off + 64, len(bad)):
        bad[i] = (bad[i] * 29 + 17) & 0xFF
    buf = (C.c_uint8 * len(bad)).from_buffer(bad)
    tiles[10].rbsp = C.cast(buf, C.POINTER(K.u8))
    out, rc, st = decoder.decode_grids([img], return_status=True)
    assert rc == K.HEIC_E_BITSTREAM
    assert st[10].code != 0
    assert all(st[t].code == 0 for t in range(48) if t != 10)
    # every other tile is still bit-exact
    for t in range(48):
        if t == 10:
            continue
        r, c = divmod(t, 8)
        y0, x0 = r * 512, c * 512
        assert np.array_equal(out[0, y0:min(y0 + 512, 3024), x0:min(x0 + 512, 4032)],
                              oracle_rgb[y0:min(y0 + 512, 3024), x0:min(x0 + 512, 4032)]), t
    # truncated stream: must terminate (zeros past the end) and be flagged
    img2, tiles2 = permuted_image(base, list(range(48)))
    tiles2[0].rbsp_len = base.tiles[0].header.slice_data_byte_offset + 20
    # keep the entry points inside the (now short) RBSP so the host accepts the descriptor
    for k in range(1, 16):
        tiles2[0].header.substream_offset[k] = min(base.tiles[0].header.substream_offset[k], 20)
    _, rc2, st2 = decoder.decode_grids([img2], return_status=True)
    assert rc2 == K.HEIC_E_BITSTREAM and st2[0].code != 0


def test_invalid_descriptors_are_rejected_on_the_host(decoder, heic_file):
    base = heic_file.primary
    img, _ = permuted_image(base, list(range(48)))
    img.n_tiles = 47
    with pytest.raises(H.HeicError) as e:
        decoder.decode_grids([img])
    assert e.value.code == K.HEIC_E_INVALID_ARG
    img, _ = permuted_image(base, list(range(48)))
    img.sps.bit_depth_luma_minus8 = 2
    with pytest.raises(H.HeicError) as e:
        decoder.decode_grids([img])
    assert e.value.code == K.HEIC_E_UNSUPPORTED


def test_aux_gain_map_monochrome_with_partial_ctbs(decoder, heic_file):
    """Fixture item 52: a single 2016x1512 hvc1 picture, 4:0:0 — not a multiple of the CTB size, no chroma."""
    aux = heic_file.aux_images
    assert len(aux) >= 1
    img = aux[0]
    assert img.sps.chroma_format_idc == 0
    td = img.tiles[0]
    ref = O.decode_picture(img.sps, img.pps, td.header, (td.rbsp, td.rbsp_len), intermediates=False)
    (y, _, _), = decoder.decode_grids_yuv([img])
    w, h = img.output_width, img.output_height
    assert y.shape == (h, w)
    assert np.array_equal(y, ref["plane"][0][:h, :w])
    rgb = decoder.decode_grids([img])[0]
    assert np.array_equal(rgb[..., 0], rgb[..., 1])  # grey


def test_colour_stitch_config2_synthetic_planes(decoder):
    """BASELINE.json configs[1]: synthetic 8-bit 4:2:0 512x512 grid tiles -> RGB + stitch, bit-exact vs the colour definition."""
    import torch

    rng = np.random.default_rng(0)
    n_tiles, tw, th = 48, 512, 512
    planes = rng.integers(0, 256, n_tiles * tw * th * 3 // 2, dtype=np.uint8)
    # corner cases in the first tile: all 27 combinations of 0/128/255
    combos = [(a, b, c) for a in (0, 128, 255) for b in (0, 128, 255) for c in (0, 128, 255)]
    for i, (yv, cbv, crv) in enumerate(combos):
        planes[(2 * (i // 16)) * tw + 2 * (i % 16)] = yv
        planes[tw * th + (i // 16) * (tw // 2) + (i % 16)] = cbv
        planes[tw * th + (tw // 2) * (th // 2) + (i // 16) * (tw // 2) + (i % 16)] = crv
    for full_range, mc in [(1, 6), (1, 1), (0, 6), (0, 1)]:
        ref = O.color_stitch(planes, 6, 8, tw, th, 4032, 3024, full_range, mc)
        d_planes = torch.from_numpy(planes).cuda()
        d_rgb = torch.zeros((3024, 4032, 3), dtype=torch.uint8, device="cuda")
        decoder.color_stitch(d_planes.data_ptr(), 1, 6, 8, tw, th, 4032, 3024, d_rgb.data_ptr(), 4032 * 3, 4032 * 3024 * 3, full_range, mc)
        torch.cuda.synchronize()
        assert np.array_equal(d_rgb.cpu().numpy(), ref), (full_range, mc)
    # ragged canvas: odd crop inside the mosaic
    ref = O.color_stitch(planes, 6, 8, tw, th, 1001, 777, 1, 6)
    d_rgb = torch.zeros((777, 1001, 3), dtype=torch.uint8, device="cuda")
    decoder.color_stitch(d_planes.data_ptr(), 1, 6, 8, tw, th, 1001, 777, d_rgb.data_ptr(), 1001 * 3, 1001 * 777 * 3, 1, 6)
    torch.cuda.synchronize()
    assert np.array_equal(d_rgb.cpu().numpy(), ref)


def test_launch_accounting(decoder, heic_file):
    n0 = decoder.launch_count()
    decoder.decode_grids([heic_file.primary])
    assert decoder.launch_count() - n0 >= 11  # cabac + list + 5 transform + intra + 2 deblock + colour (SAO applied inside it)


def test_async_submit_wait_two_jobs_in_flight(decoder, heic_file, oracle_rgb):
    """heic_b200_decode_grids_submit / _job_wait: two jobs in flight on one context, results independent and exact."""
    base = heic_file.primary
    rng = np.random.default_rng(11)
    perm = rng.permutation(48)
    img2, keep = permuted_image(base, perm)
    out_a = np.zeros((3, 3024, 4032, 3), np.uint8)
    out_b = np.zeros((2, 3024, 4032, 3), np.uint8)
    ja = decoder.submit_grids([base, base, base], out_a)
    jb = decoder.submit_grids([img2, base], out_b)
    decoder.wait_job(ja)
    decoder.wait_job(jb)
    for i in range(3):
        assert np.array_equal(out_a[i], oracle_rgb)
    assert np.array_equal(out_b[1], oracle_rgb)
    for d, s in enumerate(perm):
        r, c = divmod(d, 8)
        rs, cs = divmod(int(s), 8)
        hh = min(512, 3024 - r * 512, 3024 - rs * 512)
        ww = min(512, 4032 - c * 512, 4032 - cs * 512)
        assert np.array_equal(out_b[0, r * 512:r * 512 + hh, c * 512:c * 512 + ww], oracle_rgb[rs * 512:rs * 512 + hh, cs * 512:cs * 512 + ww])


def test_raw_nal_payloads_unescaped_on_the_gpu(decoder, heic_file, oracle_rgb):
    """heic_tile_desc::escaped = 1: the library removes the emulation prevention bytes and re-bases the entry points on
    the GPU (unescape_kernel); alone and mixed with host-unescaped images in one call."""
    out = decoder.decode_grids([heic_file.primary_raw, heic_file.primary, heic_file.primary_raw])
    for i in range(3):
        assert np.array_equal(out[i], oracle_rgb)
    # a resident batch decoded twice: the re-basing must not be applied a second time
    b = decoder.batch([heic_file.primary_raw])
    b.decode()
    b.decode()
    b.sync()
    assert all(s.code == 0 for s in b.status())
    assert np.array_equal(b.download_rgb()[0], oracle_rgb)
    b.close()


def test_gpu_unescape_matches_reference_vectors_and_host(decoder):
    """unescape_kernel against the reference's 12 emulation-prevention vectors (rbsp_reader.rs:186-303) and against the
    host restatement on random payloads dense in 00 00 03 patterns, incl. the re-basing of entry points."""
    import bisect

    from tests.test_reference_vectors import EPB_CASES

    for data, expected in EPB_CASES:
        if not data:
            continue
        got, _, _ = decoder.unescape(bytes(data))
        assert list(got) == (data if expected is None else expected), data
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 17, 255, 4096, 4097, 70001):
        # a small alphabet makes the patterns (and their corner cases: 00 00 03 03, 00 00 00 03, runs of them) frequent
        payload = rng.choice(np.array([0, 0, 0, 3, 3, 1, 4, 255], np.uint8), size=n).tobytes()
        rbsp, epb = H.remove_emulation_prevention(payload, with_positions=True)
        if len(epb) > 1024:
            continue
        data_off = min(n, 9)
        raw_sub = sorted(int(x) for x in rng.integers(0, n - data_off + 1, size=min(16, n)))
        got, off, sub = decoder.unescape(payload, data_off, raw_sub)
        assert got == rbsp, n
        removed = lambda x: bisect.bisect_left(epb, x)
        assert off == data_off - removed(data_off)
        assert sub == [data_off + s - removed(data_off + s) - off for s in raw_sub]
    # more removals in one NAL unit than the kernel's table: rejected, not mis-decoded
    with pytest.raises(H.HeicError):
        decoder.unescape(bytes([0, 0, 3]) * 2000)


def test_corrupt_tiles_in_a_multi_group_batch(decoder, heic_file, oracle_rgb):
    """Five CABAC groups (160 tiles) with corrupt tiles in different groups: under the thread-per-substream mapping the
    two persistent CTAs hand groups over warp by warp, and a failed tile must neither stall that nor leak into the next
    group handled by the same lanes."""
    base = heic_file.primary
    imgs, keep = [], []
    rng = np.random.default_rng(3)
    bad_of = {0: 5, 2: 40, 3: 17}  # image -> corrupted tile
    for i in range(4):
        img, tiles = permuted_image(base, list(range(48)))
        if i in bad_of:
            t = bad_of[i]
            raw = bytearray(bytes(base.tiles[t].rbsp[: base.tiles[t].rbsp_len]))
            off = base.tiles[t].header.slice_data_byte_offset
            for j in range(off + 40 + 13 * i, len(raw)):
                raw[j] = (raw[j] * 31 + 7 + i) & 0xFF
            buf = (C.c_uint8 * len(raw)).from_buffer(raw)
            tiles[t].rbsp = C.cast(buf, C.POINTER(K.u8))
            keep.append(buf)
        imgs.append(img)
        keep.append(tiles)
    out, rc, st = decoder.decode_grids(imgs, return_status=True)
    assert rc == K.HEIC_E_BITSTREAM
    for i in range(4):
        for t in range(48):
            if bad_of.get(i) == t:
                assert st[i * 48 + t].code != 0
                continue
            assert st[i * 48 + t].code == 0, (i, t)
            r, c = divmod(t, 8)
            y0, x0 = r * 512, c * 512
            assert np.array_equal(out[i, y0:min(y0 + 512, 3024), x0:min(x0 + 512, 4032)],
                                  oracle_rgb[y0:min(y0 + 512, 3024), x0:min(x0 + 512, 4032)]), (i, t)


@pytest.mark.gpu
def test_output_layout_is_validated_before_anything_is_queued(decoder, heic_file):
    """image i goes to rgb_out + i * image_stride, so every image must fit its slot -- checked for all images
    up front (a later, taller image used to let the D2H copy run past its slot)."""
    import ctypes as C

    from heif_b200 import _capi as K

    img = heic_file.primary
    aux = heic_file.aux_images[0]  # 2016 x 1512: a different canvas than the primary's 4032 x 3024
    arr = (K.ImageDesc * 2)()
    C.memmove(C.byref(arr, 0), C.byref(aux), C.sizeof(K.ImageDesc))
    C.memmove(C.byref(arr, C.sizeof(K.ImageDesc)), C.byref(img), C.sizeof(K.ImageDesc))
    pitch = 4032 * 3
    out = np.zeros((2, 3024, pitch), np.uint8)
    lib = K.load()
    # slots sized for the small image: the second image does not fit -> rejected, nothing written
    rc = lib.heic_b200_decode_grids(decoder._h, arr, 2, out.ctypes.data, pitch, 1512 * pitch, 0, None)
    assert rc == K.HEIC_E_INVALID_ARG and not out.any()
    # a pitch too small for the second image is caught before the first image is queued
    rc = lib.heic_b200_decode_grids(decoder._h, arr, 2, out.ctypes.data, 2016 * 3, out.strides[0], 0, None)
    assert rc == K.HEIC_E_INVALID_ARG and not out.any()
    rc = lib.heic_b200_decode_grids(decoder._h, arr, 2, out.ctypes.data, pitch, out.strides[0], 0, None)
    assert rc == 0 and out[0].any() and out[1].any()
    with pytest.raises(ValueError):
        decoder.decode_grids([aux, img])


@pytest.mark.gpu
def test_two_devices_in_one_process(built, heic_file, oracle_rgb):
    """the shared-memory opt-ins are per device; a second context on another GPU must decode (67 KB intra CTAs)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import heif_b200

    with heif_b200.HeicDecoder(device=0) as d0, heif_b200.HeicDecoder(device=1) as d1:
        assert np.array_equal(d0.decode(heic_file), oracle_rgb)
        assert np.array_equal(d1.decode(heic_file), oracle_rgb)
