"""SIMT model of the state-machine CABAC experiment (tools/experiments/cabac_fsm) on the host: one CTA of 32 columns x R row slots decodes
pool tiles as the device schedules them.  Prints, per warp-iteration, the lanes doing useful work, the lanes waiting for the
row above, and how many different state bodies a warp ran -- the three numbers the kernel's cost is made of.
    python tools/fsm_model.py [--tiles 96] [--slots 8] [--pool 64]"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import heif_b200 as H  # noqa: E402
from heif_b200 import _capi as K  # noqa: E402
from tests.synth import pool as P  # noqa: E402

STATES = ("TILE_NEXT CTU_BEGIN WAIT SAO_ML SAO_MU SAO_T1 SAO_T2 SAO_OFF SAO_SIGN SAO_BAND SAO_CLASS CQT_SPLIT PART PREV PU_MODE CHROMA1 "
          "CHROMA2 TT_SPLIT TT_CB TT_CR TT_LUMA DQP_PREFIX DQP_SIGN DQP_EG EG_PRE EG_SUF RC_TSKIP RC_LAST_PRE RC_LAST_SUF CSBF SIG DC GT1 GT2 "
          "SIGN LEVEL LEVEL_ESC LEVEL_ONE EOS EOSUB").split()

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=96)
ap.add_argument("--slots", type=int, default=8)
ap.add_argument("--cols", type=int, default=32)
ap.add_argument("--pool", type=int, default=64)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--order", default="heavy_first", choices=["heavy_first", "random", "light_first"])
args = ap.parse_args()

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")])
lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "_build", "libcabac_emul.so"))
f = H.HeicFile(open(os.path.join(ROOT, "tests", "golden", "halfmoonbay.heic"), "rb").read())
pool = P.build_pool(f, args.pool)
rng = np.random.default_rng(args.seed)
sel = rng.choice(len(pool), size=min(args.tiles, len(pool)), replace=False)
sz = pool.sizes()[sel]
if args.order == "heavy_first":
    sel = sel[np.argsort(-sz, kind="stable")]
elif args.order == "light_first":
    sel = sel[np.argsort(sz, kind="stable")]
n = len(sel)
hdrs = (C.POINTER(K.SliceHeader) * n)(*[C.pointer(pool.descs[int(s)].header) for s in sel])
rbsps = (C.c_void_p * n)(*[C.cast(pool.descs[int(s)].rbsp, C.c_void_p) for s in sel])
lens = (C.c_uint32 * n)(*[pool.descs[int(s)].rbsp_len for s in sel])
queue = (C.c_uint32 * n)(*range(n))
out = (C.c_uint64 * (8 + 2 * 64))()
rc = lib.emul_fsm_simulate(C.byref(pool.sps), C.byref(pool.pps), hdrs, rbsps, lens, n, queue, n, args.cols, args.slots, out)
assert rc == 0, rc
iters, busy, wait, dead, states, bins, engine, n_states = [int(out[i]) for i in range(8)]
lanes = iters * args.cols
print(f"{n} tiles ({args.order}), {args.cols} columns x {args.slots} slots: {bins / 1e6:.1f} Mbins, {iters} warp-iterations")
print(f"  lanes per warp-iteration: busy {busy / iters:.1f}  waiting {wait / iters:.1f}  finished {dead / iters:.1f}   (of {args.cols})")
print(f"  engine operations per warp-iteration {engine / iters:.1f};  bins per warp-iteration {bins / iters:.1f};  bins per busy lane-iteration {bins / busy:.2f}")
print(f"  different states per warp-iteration: {states / iters:.1f}")
rows = sorted(((int(out[8 + 2 * k]), int(out[9 + 2 * k]), STATES[k]) for k in range(n_states)), reverse=True)
print("  state: present in % of warp-iterations / share of lane-iterations %")
for pres, ln, name in rows[:24]:
    print(f"    {name:12s} {100 * pres / iters:5.1f}  {100 * ln / max(busy + wait, 1):5.1f}")
