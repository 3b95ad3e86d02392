"""Executed-instruction histogram by SASS opcode for every kernel of an ncu report captured with --import-source on.
    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv ; python tools/ncu_sass_hist.py sass.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

rows = csv.reader(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kernel, hdr = None, None
hist = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kernel = r[1].split("(")[0].replace("void b2r::", "").replace("b2r::", "")
        if len(r[1].split("<")) > 1:
            kernel = r[1].split("(b2r")[0].replace("void b2r::", "")
        hist.setdefault(kernel, defaultdict(lambda: [0, 0]))
        continue
    if r[0] == "Address":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or kernel is None:
        continue
    try:
        inst = int(r[hdr["Instructions Executed"]] or 0)
        thr = int(r[hdr["Thread Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        continue
    text = r[hdr["Source"]].strip()
    text = re.sub(r"^@!?U?P\d+\s+", "", text)
    op = text.split()[0] if text else "?"
    key = op.split(".")[0]
    if key in ("LDL", "STL"):
        key += " (local: spills / stack)"
    h = hist[kernel][key]
    h[0] += inst
    h[1] += thr
for k, h in hist.items():
    total = sum(v[0] for v in h.values()) or 1
    tthr = sum(v[1] for v in h.values()) or 1
    print(f"\n## {k}: {total} warp instructions executed, {tthr / total:.1f} active threads per instruction\n")
    print("| opcode | warp instructions | share | active threads / instruction |\n|---|---|---|---|")
    for op, v in sorted(h.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {op} | {v[0]} | {100 * v[0] / total:.1f} % | {v[1] / max(v[0], 1):.1f} |")
