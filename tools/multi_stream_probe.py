"""Throughput with several resident batches decoding concurrently on separate streams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import heif_b200 as H
from bench import make_images, FIXTURE, MP_PER_IMAGE

total = int(sys.argv[1]) if len(sys.argv) > 1 else 296
f = H.HeicFile(open(FIXTURE, "rb").read())
for n_streams in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]:
    per = total // n_streams
    decs = [H.HeicDecoder(0) for _ in range(n_streams)]
    keep = []
    batches = []
    for i, d in enumerate(decs):
        imgs, k = make_images(f, per, seed=1 + i)
        keep.append((imgs, k))
        batches.append(d.batch(imgs))
    for _ in range(2):
        for b in batches: b.decode()
    for b in batches: b.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 3
    for _ in range(steps):
        for b in batches: b.decode()
    for b in batches: b.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"streams={n_streams} images/stream={per}: {dt*1e3:.1f} ms per {per*n_streams} images -> {per*n_streams*MP_PER_IMAGE/dt:.0f} MP/s", flush=True)
    for b in batches: b.close()
    for d in decs: d.close()
