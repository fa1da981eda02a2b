#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: kernel, registers, spills, smem (cu++filt-demangled)."""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
rows = []
for m in re.finditer(r"Compiling entry function '([^']+)' for '(sm_\w+)'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt):
    rows.append(m.groups())
names = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.split("\n")
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    print(f"{n:70s} regs={r[5]:>3s} stack={r[2]:>4s} spill={r[3]}/{r[4]} smem={r[6]}")
