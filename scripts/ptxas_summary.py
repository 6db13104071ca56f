"""Registers / spills / shared memory per kernel from the ptxas -v logs of the last build (lib/obj/*.ptxas.txt)."""
import pathlib
import re
import subprocess
import sys

obj = pathlib.Path(__file__).resolve().parent.parent / "tornadox_b200" / "lib" / (sys.argv[2] if len(sys.argv) > 2 else "obj")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(obj.glob("*.ptxas.txt")):
    lines = f.read_text().splitlines()
    for i, ln in enumerate(lines):
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if not m:
            continue
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("tdx::", "").replace("void ", "")
        if pat and not re.search(pat, name):
            continue
        blob = " ".join(lines[i:i + 5])
        regs = re.search(r"Used (\d+) registers", blob)
        spill = re.search(r"(\d+) bytes spill stores", blob)
        print(f"{f.stem.replace('.ptxas', ''):18s} regs={regs.group(1) if regs else '?':>4s} spill={spill.group(1) if spill else '?':>4s}  {name}")
