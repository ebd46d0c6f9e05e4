"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table of per-kernel totals.

    python tools/launch_summary.py gpurun_out/launches.csv "title" "command line" > profiles/xxx.md
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for name, ns in rows:
        agg[name][0] += ns
        agg[name][1] += 1
    print(f"# {title}\n")
    print(f"Command: `{cmd}`. Times are cold-cache and serialised by the profiler: compare SHARES with bench.py's "
          f"CUDA-event breakdown, not absolutes.\n")
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total.\n")
    print("| ms | share | launches | kernel |\n|---|---|---|---|")
    for name, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {n} | `{name}` |")


if __name__ == "__main__":
    main()
