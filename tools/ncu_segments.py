"""Split the SASS of a kernel (ncu --page source --csv) into segments ending at barrier-like
instructions and report stall samples per segment, in program order."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Address" in r and "Source" in r][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seen, seq = set(), []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = r[idx["Address"]]
    if a in seen or a == "Address":
        continue
    seen.add(a)
    try:
        n = int(r[idx["# Samples"]] or 0); ex = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    st = {s.replace("stall_", ""): int(r[idx[s]] or 0) for s in stalls}
    seq.append((r[idx["Source"]].strip(), n, ex, st))
tot = sum(n for _, n, _, _ in seq)
print("total samples", tot, "instructions", len(seq))
acc, accst, first = 0, {}, 0
KEYS = ("BAR.SYNC", "BAR.ARV", "ATOMG", "EXIT", "MEMBAR", "NANOSLEEP", "WARPSYNC", "ERRBAR")
for i, (src, n, ex, st) in enumerate(seq):
    acc += n
    for k, v in st.items():
        accst[k] = accst.get(k, 0) + v
    if any(t in src for t in KEYS) or i == len(seq) - 1:
        top = [(k, v) for k, v in sorted(accst.items(), key=lambda x: -x[1])[:4] if v]
        if acc >= tot * 0.004:
            print(f"[{first:5d}-{i:5d}] {acc:6d} ({100 * acc / tot:4.1f}%) {top} | {src[:48]} ex={ex}")
        acc, accst, first = 0, {}, i + 1
