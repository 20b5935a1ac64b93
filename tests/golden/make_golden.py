"""Regenerates tests/golden/fixture_hashes.json: per-tile plane SHA-256 of halfmoonbay.heic decoded by FFmpeg's
native HEVC decoder (the independent oracle of SURVEY.md section 8(c)).  Run in the authoring container:
    python tests/golden/make_golden.py
The GPU box never runs this; the tests only read the committed JSON."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import heif_b200 as H
from oracle.ffmpeg_oracle import FFmpegHevc, annexb

f = H.HeicFile(open(os.path.join(ROOT, "tests/golden/halfmoonbay.heic"), "rb").read())
img = f.primary
ps = [f.parameter_set_nal(t) for t in (32, 33, 34)]
dec = FFmpegHevc()
tiles, allh = [], hashlib.sha256()
canvas = [np.zeros((3072, 4096), np.uint8), np.zeros((1536, 2048), np.uint8), np.zeros((1536, 2048), np.uint8)]
for t in range(img.n_tiles):
    planes = dec.decode_picture(annexb(ps + [f.tile_nal(t)]))
    tiles.append([hashlib.sha256(p.tobytes()).hexdigest() for p in planes])
    for p in planes: allh.update(p.tobytes())
    r, c = divmod(t, img.grid_cols)
    for i, p in enumerate(planes):
        s = 512 >> (1 if i else 0)
        canvas[i][r * s:(r + 1) * s, c * s:(c + 1) * s] = p
stitched = [canvas[0][:3024, :4032], canvas[1][:1512, :2016], canvas[2][:1512, :2016]]
out = {"source": "FFmpeg libavcodec 62 native hevc decoder (opencv-python-headless bundle)", "tiles": tiles,
       "all_tiles_concat": allh.hexdigest(),
       "stitched": [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in stitched],
       "stitched_mean": [round(float(p.mean()), 4) for p in stitched]}
json.dump(out, open(os.path.join(ROOT, "tests/golden/fixture_hashes.json"), "w"), indent=1)
print(out["all_tiles_concat"], out["stitched"], out["stitched_mean"])
