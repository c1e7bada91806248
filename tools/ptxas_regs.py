#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from a `-Xptxas -v` log.  usage: ptxas_regs.py build/kernels.ptxas.log [filter]"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(spb::WalkParams\)|void spb::|\(spb::\w+\)", "", name)
    if flt in name:
        print(f"{name:70s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill {m.group(3)}/{m.group(4)} smem {m.group(6)}")
