"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if hdr and r[0].isdigit() and r[2] == '-':   # a CUDA source line summary row
        d = dict(zip(hdr[4:], r[4:]))
        try:
            out.append((int(d.get("# Samples", 0) or 0), int(d.get("Instructions Executed", 0) or 0), cur, int(r[0]), r[1].strip()[:110], d))
        except ValueError:
            pass
tot_s = sum(o[0] for o in out); tot_i = sum(o[1] for o in out)
print(f"total samples {tot_s}, total warp-instructions {tot_i}")
keys = ["stall_long_sb","stall_short_sb","stall_wait","stall_branch_resolving","stall_sleep","stall_no_inst","stall_barrier","stall_lg","stall_mio","stall_math","stall_selected","stall_not_selected"]
for s, i, f, ln, src, d in sorted(out, reverse=True)[:top]:
    st = " ".join(f"{k[6:]}={d[k]}" for k in keys if k in d and d[k] not in ("0", "")) 
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins  {f}:{ln}: {src}\n        {st}")
