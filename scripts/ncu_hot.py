#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel from an ncu report (source page)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
# first kernel block only
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in body)
print(f"{len(body)} SASS lines, {tot} samples")
# cumulative by position to see loop regions
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:top]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {100*int(r[isamp])/max(tot,1):5.1f}%  ex={r[iex]:>8s}  {r[isrc].strip()[:90]}")
