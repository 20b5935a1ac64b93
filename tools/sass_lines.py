"""Static SASS instruction count per source line: python tools/sass_lines.py build/obj/x.cu.o [kernel-substring] [top]"""
import collections, os, re, subprocess, sys, tempfile
obj = os.path.abspath(sys.argv[1]); filt = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, stdout=subprocess.DEVNULL)
for cub in os.listdir(d):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    cnt = collections.Counter(); cur = None; fn = None; per_fn = collections.Counter()
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l) and fn and filt in fn:
            per_fn[fn] += 1
            if cur: cnt[cur] += 1
    for f, n in per_fn.items(): print(n, f[:100])
    for k, v in cnt.most_common(top): print(f"{v:6d}  {k[0]}:{k[1]}")
