#!/usr/bin/env python3
"""Rebuild libfosphor_b200.so with -Xptxas -v and print registers / spills / shared memory per kernel."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "gr-fosphor_b200", "build.py"), "-Xptxas", "-v"],
                     capture_output=True, text=True)
txt = out.stdout + out.stderr
names = re.findall(r"Compiling entry function '([^']+)'", txt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
blocks = txt.split("Compiling entry function")[1:]
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for name, blk in zip(dem, blocks):
    short = re.sub(r"fosphor_b200::", "", name).split("(")[0]
    if flt and flt not in short:
        continue
    used = re.search(r"Used (\d+) registers", blk)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    print("%-90s regs %s spill %s/%s" % (short[:90], used.group(1) if used else "?",
                                         spill.group(1) if spill else "?", spill.group(2) if spill else "?"))
if out.returncode:
    print(txt[-3000:])
    sys.exit(out.returncode)
