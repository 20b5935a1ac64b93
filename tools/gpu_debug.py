"""Developer aid: runs the pipeline stage by stage on the fixture and prints a mismatch summary vs the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import heif_b200 as H
from oracle import oracle_py as O

data = open(os.path.join(ROOT, "tests/golden/halfmoonbay.heic"), "rb").read()
f = H.HeicFile(data)
img = f.primary
tiles = list(range(img.n_tiles)) if len(sys.argv) < 2 else [int(x) for x in sys.argv[1].split(",")]
t0 = time.time()
refs = {t: O.decode_picture(img.sps, img.pps, img.tiles[t].header, (img.tiles[t].rbsp, img.tiles[t].rbsp_len)) for t in tiles}
print("oracle decode", time.time() - t0, "s")
dec = H.HeicDecoder(0)
b = dec.batch([img])

def cmp(name, got, ref, t):
    if got.shape != ref.shape:
        print(f"  tile {t} {name}: shape {got.shape} vs {ref.shape}"); return False
    bad = np.flatnonzero(got.ravel() != ref.ravel())
    if bad.size:
        print(f"  tile {t} {name}: {bad.size} mismatches, first at {bad[:6]} got {got.ravel()[bad[:6]]} ref {ref.ravel()[bad[:6]]}")
        return False
    return True

stages = [("cabac", H.STAGE_CABAC), ("transform", H.STAGE_TRANSFORM), ("intra", H.STAGE_INTRA), ("deblock", H.STAGE_DEBLOCK), ("sao", H.STAGE_SAO)]
for name, mask in stages:
    t0 = time.time(); b.run(mask); b.sync(); dt = time.time() - t0
    ok = True
    if name == "cabac":
        st = b.status()
        for t in tiles:
            if st[t].code != 0 or st[t].bins_decoded != refs[t]["bins"] or st[t].ctus_decoded != refs[t]["ctus"]:
                print(f"  tile {t} status code={st[t].code} bins={st[t].bins_decoded}/{refs[t]['bins']} ctus={st[t].ctus_decoded}/{refs[t]['ctus']}"); ok = False
    for t in tiles:
        d, r = b.dump_tile(t), refs[t]
        if name == "cabac":
            ok &= cmp("tu_map", d["tu_map"], r["tu_map"], t)
            for c in range(3): ok &= cmp(f"level{c}", d["coeff"][c], r["level"][c], t)
            ok &= cmp("qp", d["qp_map"], r["qp_map"], t); ok &= cmp("sao", d["sao"], r["sao"], t)
        elif name == "transform":
            for c in range(3): ok &= cmp(f"resid{c}", d["coeff"][c], r["resid"][c], t)
        else:
            key = {"intra": "recon", "deblock": "deblocked", "sao": "plane"}[name]
            for c in range(3): ok &= cmp(f"{key}{c}", d["plane"][c], r[key][c], t)
    print(f"stage {name}: {'OK' if ok else 'MISMATCH'}  ({dt*1e3:.2f} ms incl. launch+sync)")
b.run(H.STAGE_COLOR); b.sync()
rgb = b.download_rgb()[0]
planes = np.concatenate([np.concatenate([p.ravel() for p in refs[t]["plane"]]) for t in range(img.n_tiles)]) if len(tiles) == img.n_tiles else None
if planes is not None:
    ref = O.color_stitch(planes, img.grid_rows, img.grid_cols, 512, 512, img.output_width, img.output_height, img.sps.video_full_range_flag, img.sps.matrix_coeffs)
    print("color+stitch:", "OK" if np.array_equal(rgb, ref) else f"MISMATCH {(rgb != ref).sum()}")
# whole-pipeline timing, resident
for it in range(3):
    t0 = time.time(); b.decode(); b.sync(); print(f"full decode (1 image resident): {(time.time()-t0)*1e3:.2f} ms")
print("launches", dec.launch_count())
