#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (ncu launch list) and a --set full report into compact tables."""
import collections
import csv
import subprocess
import sys


def launch_list(path, top=16):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(row["Kernel Name"][:60], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:9.1f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.2f}  {k}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]


def full(path, out_csv=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index("Kernel Name")
    table = [["kernel"] + [w for w, _ in cols], [""] + [units[i] for _, i in cols]]
    for r in rows[2:]:
        table.append([r[ki][:40]] + [r[i] for _, i in cols])
    for r in table[2:]:
        print(r[0])
        for (w, _), v, u in zip(cols, r[1:], table[1][1:]):
            print(f"    {w:68s} {v} {u}")
    if out_csv:
        csv.writer(open(out_csv, "w")).writerows(table)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launch_list(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
