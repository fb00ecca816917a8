"""Registers / spills per kernel from an `nvcc -Xptxas -v` log (developer tool).  usage: ptxas_summary.py build_ab/x.log [filter]"""
import re, subprocess, sys
log = open(sys.argv[1]).read().split("\n")
flt = sys.argv[2] if len(sys.argv) > 2 else ""
name = None
for i, l in enumerate(log):
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
    if m and name:
        sp = m.groups()
    m = re.search(r"Used (\d+) registers", l)
    if m and name:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void mrf::", "")
        if flt in dem:
            print(f"{dem:60s} regs {m.group(1):>3s}  stack {sp[0]:>4s}  spill st/ld {sp[1]}/{sp[2]}")
        name = None
