"""Scratch: attribute an ncu source-page export (SASS rows) to CUDA source lines.

    ncu -i X.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv <cubin> <kernel-name-substring> [top]

The per-instruction rows of the ncu export are joined, in order, with `nvdisasm -g` of the same kernel (the .so must be
the one that was profiled), whose `//## File "...", line N` markers name the innermost source line of every instruction."""
import csv
import collections
import re
import subprocess
import sys

src_csv, cubin, pattern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []      # (file, line) per instruction, in order
inside = False
cur = ("?", 0)
for ln in text:
    if ln.startswith("\t.section\t.text."):
        inside = pattern in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
        lines.append(cur)

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
assert len(body) == len(lines), (len(body), len(lines))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
tot_samples = tot_inst = 0
for r, key in zip(body, lines):
    a = agg[key]
    s = int(r[col["# Samples"]]); i = int(r[col["Instructions Executed"]])
    a["samples"] += s; a["inst"] += i
    tot_samples += s; tot_inst += i
    for h in stall_cols:
        a[h] += int(r[col[h]])
print(f"total samples {tot_samples}, warp instructions {tot_inst}")
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s}  top stalls")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((a[h], h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{key[0] + ':' + str(key[1]):28s} {100 * a['samples'] / tot_samples:6.2f} {100 * a['inst'] / tot_inst:6.2f}  " +
          " ".join(f"{n}:{100 * v / max(1, a['samples']):.0f}%" for v, n in st))
