"""Short decode of the bench workload (pool batch) for ncu: builds the batch, runs --decodes full
decodes (or only --stages), prints per-stage CUDA-event times.  `ncu -k regex:<kernel> -s <skip> -c 1 python tools/profile_batch.py`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import heif_b200 as H  # noqa: E402
from bench import FIXTURE, SEED, check_groups_distinct  # noqa: E402
from tests.synth import pool as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=592)
ap.add_argument("--pool", type=int, default=512)
ap.add_argument("--decodes", type=int, default=2)
ap.add_argument("--stage-reps", type=int, default=1)
args = ap.parse_args()

f = H.HeicFile(open(FIXTURE, "rb").read())
pool = P.build_pool(f, args.pool)
images, keep, ids = P.compose_images(pool, f.primary, range(args.batch), SEED)
dec = H.HeicDecoder(0)
b = dec.batch(images)
try:
    print("groups:", check_groups_distinct(b, ids.reshape(-1)), flush=True)
except SystemExit as e:  # e.g. HEIC_B200_CABAC_PLAIN_SORT=1: measured anyway, but said
    print("groups:", e, flush=True)
s = torch.cuda.ExternalStream(b.stream)
for _ in range(args.decodes):
    b.decode()
b.sync()
st = b.status()
assert all(st[i].code == 0 for i in range(b.n_tiles))
bins = sum(st[i].bins_decoded for i in range(b.n_tiles))
names = [("cabac", H.STAGE_CABAC), ("transform", H.STAGE_TRANSFORM), ("intra", H.STAGE_INTRA), ("deblock", H.STAGE_DEBLOCK),
         ("sao+color", H.STAGE_SAO | H.STAGE_COLOR)]
for _ in range(args.stage_reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record(s)
    for i, (n, m) in enumerate(names):
        b.run(m)
        ev[i + 1].record(s)
    b.sync()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]
    print("  ".join(f"{n} {t:.2f} ms" for (n, _), t in zip(names, ms)), f" total {sum(ms):.2f} ms  bins {bins / 1e9:.3f} G "
          f"({bins / ms[0] / 1e6 / 148:.3f} Gbin/s/SM)", flush=True)
b.close()
dec.close()
