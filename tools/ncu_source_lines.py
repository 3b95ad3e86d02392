"""Per-source-line instruction and stall-sample shares from an ncu report captured with --import-source on.
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_lines.py src.csv <kernel name substring> [top]
Launches of the same kernel are summed."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file, cur_fn, hdr = "?", "", None
agg = defaultdict(lambda: [0, 0, ""])  # (file, line) -> [inst, samples, text]
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or want not in cur_fn:
        continue
    if r[0]:  # a CUDA source line: aggregated metrics of its SASS
        try:  # a source line holding inline asm with quotes breaks the csv fields: skip it
            key = (cur_file, int(r[0]))
            inst, smp = int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["# Samples"]] or 0)
        except ValueError:
            continue
        agg[key][0] += inst
        agg[key][1] += smp
        agg[key][2] = r[1].strip()
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"kernel *{want}*: {ti} warp instructions, {ts} samples, {len(agg)} source lines")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * v[1] / ts:5.1f}% smp {100 * v[0] / ti:5.1f}% inst  {f}:{l:<5d} {v[2][:110]}")
