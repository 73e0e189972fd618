#!/usr/bin/env python3
"""Static SASS instruction census of one kernel by source line (needs -lineinfo): where the body's bytes come from.
usage: sass_census.py <kernel substring> [top N] [lib.so]"""
import collections, os, re, subprocess, sys, tempfile
pat = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = os.path.abspath(sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "akari_render_b200", "libakari_b200.so"))
root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin") and "scene_build" not in f][0]
lines = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and pat in l][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//-----")), len(lines))
cur, cnt, fcnt = None, collections.Counter(), collections.Counter()
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[cur] += 1
        fcnt[cur[0] if cur else None] += 1
def src(f, n):
    for sub in ("akari_render_b200/csrc/device", "akari_render_b200/csrc"):
        p = os.path.join(root, sub, f)
        if os.path.exists(p):
            return open(p).read().split("\n")[n - 1].strip()[:110]
    return ""
print(sum(cnt.values()), "instructions;", dict(fcnt.most_common(8)))
for k, v in cnt.most_common(top):
    print(f"{v:5d}  {k[0]}:{k[1]}  {src(*k)}" if k else f"{v:5d}  ?")
