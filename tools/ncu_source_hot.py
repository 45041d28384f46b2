"""Aggregate an `ncu --page source --csv` dump: top SASS instructions by stall samples, with reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    tot += n
    data.append((n, r))
data.sort(key=lambda x: -x[0])
print("total samples", tot)
agg = {s: 0 for s in stalls}
for n, r in data:
    for s in stalls:
        try:
            agg[s] += int(r[idx[s]] or 0)
        except ValueError:
            pass
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot * 0.01})
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for n, r in data[:topn]:
    rs = {s.replace("stall_", ""): int(r[idx[s]] or 0) for s in stalls if (r[idx[s]] or "0") != "0"}
    rs = dict(sorted(rs.items(), key=lambda x: -x[1])[:3])
    print(f"{n:6d} {100*n/tot:5.1f}%  {r[idx['Source']][:70]:70s} {rs}")
