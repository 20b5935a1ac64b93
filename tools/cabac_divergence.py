"""How much of the CABAC throughput of bench.py comes from warps whose 32 lanes decode copies of the same tile?
bench.py's batch repeats the fixture's 48 tiles, and sorting by slice size puts identical tiles into the same warp.
This probe decodes synthetic 512x512 WPP pictures (random syntax: the worst case for lane divergence) two ways:
  same     : 32 distinct pictures, 32 copies each  -> every warp holds 32 copies of one picture
  distinct : 1024 distinct pictures (same config)  -> every warp holds 32 different pictures of similar size
and prints CABAC bins/s for both.   python tools/cabac_divergence.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import heif_b200 as H
from tests.synth import synth

CFG = dict(width=512, height=512, log2_ctb=5, wpp=1, init_qp_minus26=-10, lps_gain=1.5)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dec = H.HeicDecoder(0)

def run(pics):
    b = dec.batch([p.desc for p in pics])
    st = torch.cuda.ExternalStream(b.stream)
    for _ in range(2):
        b.run(H.STAGE_CABAC)
    b.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(st)
    for _ in range(reps):
        b.run(H.STAGE_CABAC)
    e1.record(st)
    b.sync()
    s = b.status()
    assert all(s[i].code == 0 for i in range(b.n_tiles))
    bins = sum(s[i].bins_decoded for i in range(b.n_tiles))
    ms = e0.elapsed_time(e1) / reps
    b.close()
    return ms, bins

t0 = time.time()
distinct = [synth.encode(1000 + i, **CFG) for i in range(N)]
print(f"{N} pictures encoded in {time.time() - t0:.0f} s, {np.mean([len(p.slice_nal) for p in distinct]):.0f} B each on average")
same = [distinct[i // 32] for i in range(N)]
for name, pics in (("same", same), ("distinct", distinct)):
    ms, bins = run(pics)
    print(f"{name:9s} {ms:8.3f} ms  {bins / ms / 1e6:8.2f} Gbin/s  ({bins / N:.0f} bins per picture)")
