#!/usr/bin/env python
"""Bucket ncu SASS-page stall samples between barriers/cp.async waits: python tools/sass_phases.py sass.csv [kernel-substr]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
# split per kernel
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if want not in b["name"]:
        continue
    hdr, data = b["rows"][0], b["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    S, X = ix["# Samples"], ix["Instructions Executed"]
    tot = sum(int(r[S]) for r in data)
    print(b["name"][:90], "samples", tot)
    marks = [i for i, r in enumerate(data) if "BAR.SYNC" in r[1] or "DEPBAR" in r[1]]
    cuts = [0] + marks + [len(data)]
    for a, c in zip(cuts[:-1], cuts[1:]):
        n = sum(int(r[S]) for r in data[a:c])
        if n * 200 < tot:
            continue
        st = {}
        for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_lg", "stall_mio", "stall_not_selected", "stall_selected"):
            st[k[6:]] = sum(int(r[ix[k]]) for r in data[a:c])
        top = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        print(f"  sass[{a:5d},{c:5d}) {data[a][1].strip()[:28]:28s} samples {n:6d} ({100*n/tot:4.1f}%) inst {sum(int(r[X]) for r in data[a:c]):10d}  {top}")
