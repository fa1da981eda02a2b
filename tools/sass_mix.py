#!/usr/bin/env python
"""Static SASS instruction mix per kernel of an object/cubin (cuobjdump -sass): FP64 vs everything else."""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else ".")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
names = [f.split("\n")[0].strip() for f in funcs]
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout.split("\n") if names else []
for f, n in sorted(zip(funcs, dem), key=lambda t: t[1]):
    n = n.replace("void ", "").replace("(int)", "").replace("(bool)", "")
    n = re.sub(r">\(.*", ">", n) if "<" in n else re.sub(r"\(.*", "", n)
    if not re.search(pat, n): continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m: ops[m.group(2).split(".")[0]] += 1
    tot = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    mv = ops["MOV"] + ops["IMAD"] + ops["IADD3"] + ops["LEA"] + ops["UMOV"]
    print(f"{n:42s} total={tot:5d} fp64={fp64:5d} ({100*fp64/max(tot,1):4.1f}%) DFMA={ops['DFMA']:4d} DMUL={ops['DMUL']:4d} DADD={ops['DADD']:4d} "
          f"MUFU={ops['MUFU']:3d} LDG={ops['LDG']:3d} STG={ops['STG']:3d} BRA={ops['BRA']:3d} CALL={ops['CALL']:2d} int/mov={mv:4d} LDL/STL={ops['LDL']+ops['STL']}")
