"""Small decode set for compute-sanitizer: the fixture image + a few synthetic geometries, through the C ABI."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import heif_b200 as H
from tests.synth import synth

dec = H.HeicDecoder(0)
f = H.HeicFile(open(os.path.join(ROOT, "tests/golden/halfmoonbay.heic"), "rb").read())
rgb = dec.decode(f)
print("fixture", rgb.shape, int(rgb.sum()))
pics = [synth.encode(0, **cfg) for cfg in (dict(), dict(wpp=0), dict(log2_ctb=4, log2_max_tb=4, width=72, height=104),
                                            dict(log2_ctb=6, width=136, height=200, wpp=0), dict(chroma_format_idc=0),
                                            dict(width=8, height=8), dict(transform_skip=1, sign_data_hiding=1, cu_qp_delta=1, diff_cu_qp_delta_depth=2))]
res = dec.decode_grids_yuv([p.desc for p in pics])
print("synthetic", [int(r[0].sum()) for r in res])
for p in pics:
    out = dec.decode_grids([p.desc])
aux = f.aux_images[0]
print("aux", dec.decode_grids_yuv([aux])[0][0].shape)
