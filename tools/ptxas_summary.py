#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: kernel, registers, spills, smem (cu++filt-demangled)."""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
rows = []
for m in re.finditer(r"Compiling entry function '([^']+)' for '(sm_\w+)'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt):
    rows.append(m.groups())
names = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout.split("\n") if rows else []
for r, n in zip(rows, names):
    n = n.replace("void ", "").replace("(int)", "").replace("(bool)", "")
    n = re.sub(r">\(.*", ">", n) if "<" in n else re.sub(r"\(.*", "", n)
    print(f"{n:70s} regs={r[5]:>3s} stack={r[2]:>4s} spill={r[3]}/{r[4]} smem={r[6]}")
