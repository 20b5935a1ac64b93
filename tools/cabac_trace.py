"""Digest of the CABAC group timeline printed by a -DHEIC_CABAC_TRACE=1 build:
  HEIC_NVCC_EXTRA=-DHEIC_CABAC_TRACE=1 python -c "import __graft_entry__ as g; g.build_library(variant='trace')"
  HEIC_B200_LIB=heif_b200/variants/libheic_trace.so python tools/profile_batch.py --decodes 1 --stage-reps 0 > trace.log
  python tools/cabac_trace.py trace.log
Lines are `CT cta group use slot t_begin t_end bins_of_the_warp slice_bytes_of_lane0` (globaltimer ns)."""
import collections
import sys

import numpy as np

a = np.array([l.split()[1:] for l in open(sys.argv[1]) if l.startswith("CT ")], dtype=np.int64)
t0 = a[:, 4].min()
a[:, 4] -= t0
a[:, 5] -= t0
print(f"kernel span {a[:, 5].max() / 1e6:.1f} ms, {len(a)} (group, warp) records")
g = {}
for r in a:
    d = g.setdefault(int(r[1]), [1 << 62, 0, 0, int(r[7]), int(r[0]), int(r[2])])
    d[0] = min(d[0], r[4]); d[1] = max(d[1], r[5]); d[2] += r[6]
ids = sorted(g)
print("group use cta  start_ms  end_ms  dur_ms  Mbins  lane0_bytes  Mbin/s")
for i in sorted(set([0, 1, 2, 10, 50, 100, 200, 300, 443, 444, 500, 600, 700, 800, 900, len(ids) - 1])):
    if i >= len(ids):
        continue
    s, e, b, bl, cta, use = g[ids[i]]
    print(f"{ids[i]:5d} {use:3d} {cta:4d} {s / 1e6:8.2f} {e / 1e6:8.2f} {(e - s) / 1e6:7.2f} {b / 1e6:7.2f} {bl:8d} {b / max(e - s, 1) * 1e3:8.1f}")
ce = collections.defaultdict(int)
we = collections.defaultdict(int)
for r in a:
    ce[int(r[0])] = max(ce[int(r[0])], r[5])
    we[(int(r[0]), int(r[3]))] = max(we[(int(r[0]), int(r[3]))], r[5])
e = np.sort(np.array(list(ce.values()))) / 1e6
w = np.array(list(we.values())) / 1e6
print(f"CTA end times: min {e[0]:.1f} p10 {e[len(e) // 10]:.1f} p50 {e[len(e) // 2]:.1f} p90 {e[9 * len(e) // 10]:.1f} max {e[-1]:.1f} ms")
print(f"warps gone before the kernel's end: {100 * (1 - w.mean() / e[-1]):.1f} % of the warp-time")
print("groups per CTA:", dict(collections.Counter(collections.Counter(v[4] for v in g.values()).values())))
