#!/usr/bin/env python
"""Opcode mix of the hottest loop body of a kernel (cuobjdump -sass): the body of the backward branch that spans the most
instructions.  usage: sass_loop.py obj.o 'demangled-name-regex'"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n")[0].strip()
    dn = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    dn = dn.replace("(int)", "").replace("(bool)", "")
    if not re.search(pat, dn): continue
    ins = []
    for line in f.split("\n"):
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    best = None
    for a, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < a:
                span = a - int(t.group(1), 16)
                if best is None or span > best[0]: best = (span, int(t.group(1), 16), a)
    if not best: print(dn[:80], "no loop"); continue
    body = [x for x in ins if best[1] <= x[0] <= best[2]]
    c = collections.Counter(op.split(".")[0] for _, op, _ in body)
    fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"{re.sub(r'[(].*', '', dn)[:70]}: loop body {len(body)} instr, fp64 {fp64} (2*fp64+other = {fp64 + len(body)})")
    print("   " + " ".join(f"{k}:{v}" for k, v in c.most_common(24)))
