"""Print the most-sampled SASS instructions of an `ncu --page source --csv` export with their top stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for h in stalls:
        agg[h] = agg.get(h, 0) + int(r[idx[h]])
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    s = {h[6:]: int(r[idx[h]]) for h in stalls if int(r[idx[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[idx['# Samples']]):6d} {100 * int(r[idx['# Samples']]) / tot:5.1f}% exec={r[idx['Instructions Executed']]:>8s} "
          f"{r[idx['Source']].strip()[:60]:60s} {s}")
