#!/usr/bin/env python
"""Per-instruction stall summary of an ncu report (needs --import-source on / -lineinfo):

    python tools/ncu_src.py gpurun_out/X.ncu-rep [kernel-substring] [top-N]

Prints, per kernel: total samples, the stall-reason histogram and the top-N SASS instructions by stall samples.
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for b in blocks:
    if filt and filt not in b["name"]:
        continue
    h = b["hdr"]
    si = h.index("# Samples")
    ie = h.index("Instructions Executed")
    stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in b["rows"])
    print(f"\n=== {b['name'][:140]}\n    instructions {len(b['rows'])}, samples {tot}")
    hist = {h[i]: sum(int(r[i]) for r in b["rows"]) for i in stalls}
    print("    stalls:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(hist.items(), key=lambda kv: -kv[1]) if v))
    rows = sorted(enumerate(b["rows"]), key=lambda ir: -int(ir[1][si]))[:topn]
    for idx, r in sorted(rows):
        top = sorted(((int(r[i]), h[i][6:]) for i in stalls), reverse=True)[:2]
        print(f"    #{idx:5d} {100 * int(r[si]) / max(tot, 1):5.1f}%  exec {r[ie]:>9}  {r[1].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
