"""Join an ncu source-page CSV (SASS level, `ncu -i rep --page source --csv`) with the line table of
the cubin (`nvdisasm -g -c`) and print the hottest source lines by warp-stall samples.
usage: python tools/ncu_lines.py <rep.ncu-rep> <lib.so> <kernel-substring> [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, isamp, isrc = H.index("Address"), H.index("# Samples"), H.index("Source")
istall = H.index("Warp Stall Sampling (All Samples)")
samples = {}
for r in rows[hdr + 1:]:
    if len(r) == len(H):
        samples[int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])] = (int(r[isamp] or 0), r[isrc])
base = min(samples)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            infn = kname in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
agg, tot = {}, 0
for addr, (n, sass) in samples.items():
    key = line_of.get(addr - base)
    agg.setdefault(key, [0, {}])
    agg[key][0] += n
    op = sass.split()[0] if sass.split() else "?"
    if op.startswith("@"):
        op = sass.split()[1]
    agg[key][1][op] = agg[key][1].get(op, 0) + n
    tot += n
print(f"total samples {tot}")
for key, (n, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    best = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print(f"{100 * n / tot:6.2f}%  {key}  " + " ".join(f"{o}:{c}" for o, c in best))
