"""Digest the round's ncu artefacts into profiles/:  python tools/profile_digest.py <launches.csv> <ncu_full_summary.txt> <n_images> [first_n_launches] [prefix]
 - profiles/<prefix>_launches_summary.txt : per-kernel mean launch time and share of one decode
 - profiles/dram_traffic_per_image.json    : dram__bytes_read.sum + dram__bytes_write.sum per stage and image (bench.py `traffic`)"""
import collections, csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, full, n_img = sys.argv[1], sys.argv[2], int(sys.argv[3])
first = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # only the first `first` launches (the resident-batch decodes)
prefix = sys.argv[5] if len(sys.argv) > 5 else "r02_final"  # profiles/<prefix>_launches_summary.txt

rows = list(csv.reader(open(launches)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
order, t = [], collections.defaultdict(list)
for r in rows[h + 1:]:
    if len(r) < len(H):
        continue
    rec = dict(zip(H, r))
    if rec["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void\s+", "", rec["Kernel Name"])
    name = re.sub(r"\(.*", "", name).replace("heic::dev::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    v = float(rec["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(rec["Metric Unit"], 1e-6)
    if first and sum(len(x) for x in t.values()) >= first:
        break
    if name not in t:
        order.append(name)
    t[name].append(v)
n_dec = min(len(v) for v in t.values())
per_dec = sum(sum(v) / len(v) * (len(v) // n_dec) for v in t.values())
with open(os.path.join(ROOT, "profiles", prefix + "_launches_summary.txt"), "w") as f:
    f.write(f"# ncu launch list, `python bench.py --steps 2 --warmup 3 --no-cpu` (batch {n_img} images/GPU), first {n_dec} full decodes\n")
    f.write("# gpu__time_duration.sum per launch, --clock-control none; serialised + cold cache: compare SHARES with bench.py's `stages`, not absolutes\n")
    f.write(f"# {per_dec:.1f} ms per decode under ncu\n\n")
    for k in order:
        m = sum(t[k]) / len(t[k])
        f.write(f"{m:9.3f} ms/launch  {100 * m * (len(t[k]) // n_dec) / per_dec:5.1f}%  x {len(t[k])}  {k}\n")

unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
stage_of = [("cabac", "cabac"), ("tu_list", "transform"), ("transform", "transform"), ("intra", "intra"), ("deblock", "deblock"),
            ("sao", "sao"), ("color", "color_stitch")]
traffic, missing = collections.defaultdict(float), []
for b in open(full).read().split("----"):
    m = re.search(r"Kernel Name\s+(.*)", b)
    if not m:
        continue
    stage = next((s for key, s in stage_of if key in m.group(1)), None)
    tot = 0.0
    for metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        mm = re.search(metric + r"\s+([-\w.]+)\s+(\w+)", b)
        try:
            v = float(mm.group(1)) * unit.get(mm.group(2), 1)
        except (AttributeError, ValueError):
            v = float("nan")
        tot += v
    if tot != tot:
        missing.append(m.group(1).strip()[:50])
    else:
        traffic[stage] += tot
out = {"_source": f"ncu --set full, profiles/{prefix}_ncu_full_summary.txt (tools/profile_batch.py --decodes 1, batch {n_img}): "
                  f"dram__bytes_read.sum + dram__bytes_write.sum per launch / {n_img} images, summed per stage",
       "_not_captured": missing}
for s, v in traffic.items():
    out[s] = v / n_img
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic_per_image.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", prefix + "_launches_summary.txt")).read())
print(json.dumps(out, indent=1))
