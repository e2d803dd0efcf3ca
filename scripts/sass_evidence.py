#!/usr/bin/env python
"""Count the Blackwell-native SASS mnemonics per kernel of the built library (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk (1-D bulk copy),
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0), REDUX = redux.sync (CREDUX on sm_100),
F*2 = packed fp32 pairs (FADD2 / FMUL2), UBLKPF = cp.async.bulk.prefetch.L2.  Writes a table to stdout."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "depthg_b200/libdepthg_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WANT = [("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
        ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("UTCBAR", r"\bUTCBAR\b"), ("SYNCS", r"\bSYNCS\b"), ("HMMA", r"\bHMMA\b"),
        ("REDUX", r"\bC?REDUX\b"), ("LDGSTS", r"\bLDGSTS\b"), ("F*2 (packed fp32)", r"\bF(ADD|MUL|FMA)2\b"),
        ("UBLKPF", r"\bUBLKPF\b")]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("dg::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    counts[cur]["instr"] += bool(re.search(r"/\*[0-9a-f]{4,}\*/", line))
    for key, pat in WANT:
        if re.search(pat, line):
            counts[cur][key] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print("| kernel | instr | " + " | ".join(k for k, _ in WANT) + " |")
print("|---|---|" + "---|" * len(WANT))
for k, c in counts.items():
    print(f"| `{k}` | {c['instr']} | " + " | ".join(str(c[key]) for key, _ in WANT) + " |")
