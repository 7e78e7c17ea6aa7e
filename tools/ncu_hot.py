#!/usr/bin/env python
"""Hot SASS instructions of one kernel in an ncu report's source page (warp-stall samples).
   ncu -i rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv '<0>' [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = b["hdr"]
    si, ai = h.index("# Samples"), h.index("Source")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[si]) for r in b["rows"])
    print(b["name"], "total samples", tot, "instructions", len(b["rows"]))
    agg = {}
    for r in b["rows"]:
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    idx = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][si]))[:top]
    for i in sorted(idx):
        r = b["rows"][i]
        st = sorted(((int(r[c]), h[c]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:6d} {int(r[si]):7d} {100*int(r[si])/tot:5.1f}%  {r[ai].strip()[:70]:70s} {st}")
    break
