"""CABAC stage time for batches made of one tile repeated (light / medium / heavy), to separate latency from throughput."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import heif_b200 as H
from heif_b200 import _capi as K
from bench import FIXTURE

f = H.HeicFile(open(FIXTURE, "rb").read())
base = f.primary
sizes = sorted((base.tiles[t].rbsp_len, t) for t in range(48))
print("tile sizes:", [s for s, _ in sizes])
dec = H.HeicDecoder(0)

def run(tile_ids, n_images, label):
    keep, imgs = [], []
    for i in range(n_images):
        tiles = (K.TileDesc * 48)()
        for d in range(48):
            C.memmove(C.byref(tiles, d * C.sizeof(K.TileDesc)), C.byref(base.tiles[tile_ids[(i * 48 + d) % len(tile_ids)]]), C.sizeof(K.TileDesc))
        im = K.ImageDesc(); C.memmove(C.byref(im), C.byref(base), C.sizeof(K.ImageDesc)); im.tiles = C.cast(tiles, C.POINTER(K.TileDesc))
        keep.append(tiles); imgs.append(im)
    b = dec.batch(imgs)
    s = torch.cuda.ExternalStream(b.stream)
    for _ in range(2): b.run(H.STAGE_CABAC)
    b.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); b.run(H.STAGE_CABAC); e1.record(s); b.sync()
    st = b.status(); bins = sum(st[i].bins_decoded for i in range(b.n_tiles))
    ms = e0.elapsed_time(e1)
    print(f"{label:28s} images={n_images:4d} tiles={b.n_tiles:6d}  cabac {ms:8.2f} ms  {bins/1e6:9.1f} Mbins  {bins/ms/1e6:7.2f} Gbins/s", flush=True)
    b.close()

light = [t for _, t in sizes[:4]]; heavy = [t for _, t in sizes[-4:]]; mid = [t for _, t in sizes[22:26]]
for n in (37, 148, 296):
    run(light, n, "4 lightest tiles"); run(mid, n, "4 median tiles"); run(heavy, n, "4 heaviest tiles"); run(list(range(48)), n, "all 48 tiles")
