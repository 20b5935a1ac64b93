"""Small resident-batch decode for ncu: python tools/profile_run.py [n_images] [n_decodes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heif_b200 as H
from bench import make_images, FIXTURE

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
f = H.HeicFile(open(FIXTURE, "rb").read())
dec = H.HeicDecoder(0)
images, keep = make_images(f, n, seed=1)
b = dec.batch(images)
for _ in range(reps):
    b.decode()
b.sync()
st = b.status()
assert all(st[i].code == 0 for i in range(b.n_tiles))
print("ok", n, "images", dec.launch_count(), "launches")
