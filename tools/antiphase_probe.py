"""Two resident batches decoded in anti-phase: CABAC of one overlaps the remaining stages of the other."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import heif_b200 as H
from bench import make_images, FIXTURE, MP_PER_IMAGE

per = int(sys.argv[1]) if len(sys.argv) > 1 else 296
use_events = int(sys.argv[2]) if len(sys.argv) > 2 else 1
f = H.HeicFile(open(FIXTURE, "rb").read())
decs = [H.HeicDecoder(0), H.HeicDecoder(0)]
keep, batches = [], []
for i, d in enumerate(decs):
    imgs, k = make_images(f, per, seed=1 + i); keep.append((imgs, k)); batches.append(d.batch(imgs))
A, B = batches
sA, sB = (torch.cuda.ExternalStream(b.stream) for b in batches)
REST = H.STAGE_ALL & ~H.STAGE_CABAC
for b in batches: b.decode()
for b in batches: b.sync()
# serial reference
t0 = time.perf_counter()
for _ in range(2):
    for b in batches: b.decode(); 
for b in batches: b.sync()
ser = (time.perf_counter() - t0) / 2
print(f"serial: {ser*1e3:.1f} ms per {2*per} images -> {2*per*MP_PER_IMAGE/ser:.0f} MP/s")
# anti-phase: prime B with a CABAC so that its REST is valid
B.run(H.STAGE_CABAC); B.sync()
steps = 4
torch.cuda.synchronize()
t0 = time.perf_counter()
evB = None
for k in range(steps):
    if use_events and evB is not None: sA.wait_event(evB)
    A.run(H.STAGE_CABAC)
    evA = torch.cuda.Event(); evA.record(sA)
    B.run(REST)
    if use_events: sB.wait_event(evA)
    B.run(H.STAGE_CABAC)
    evB = torch.cuda.Event(); evB.record(sB)
    A.run(REST)
for b in batches: b.sync()
dt = (time.perf_counter() - t0) / steps
print(f"anti-phase (events={use_events}): {dt*1e3:.1f} ms per {2*per} images -> {2*per*MP_PER_IMAGE/dt:.0f} MP/s")
st = A.status(); assert all(st[i].code == 0 for i in range(A.n_tiles))
