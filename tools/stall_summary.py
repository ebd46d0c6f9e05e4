"""Summarise an ncu source-page CSV: instruction mix, issue utilisation and which barrier waits dominate."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_s = sum(int(r[isamp]) for r in data)
tot_e = sum(int(r[iex]) for r in data)
print(f"samples {tot_s}  warp-instructions {tot_e / 1e6:.1f}M")
waits = collections.Counter()
ops = collections.Counter()
for r in data:
    src = r[isrc].strip()
    toks = [t for t in src.split() if not t.startswith("@")]
    if toks:
        ops[toks[0].split(".")[0]] += int(r[iex])
    if "SYNCS.PHASECHK" in src or "BAR.SYNC" in src or "WARPSYNC" in src or "UTMACMDFLUSH" in src or "DEPBAR" in src:
        waits[src[:80]] += int(r[isamp])
print("top wait sites (samples):")
for k, v in waits.most_common(12):
    print(f"  {v:6d} {100 * v / tot_s:5.1f}%  {k}")
print("top sampled instructions:")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
    print(f"  {int(r[isamp]):6d} {100 * int(r[isamp]) / tot_s:5.1f}%  ex={r[iex]:>9s}  {r[isrc].strip()[:80]}")
print("opcode mix:", ", ".join(f"{k} {100 * v / tot_e:.1f}%" for k, v in ops.most_common(14)))
