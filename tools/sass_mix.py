"""Instruction mix per kernel from cuobjdump -sass (static counts)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "ilqr.jl_b200/libilqr_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "two_link"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    print(name[-60:], "total", sum(ops.values()), "fp64", ops["DFMA"] + ops["DMUL"] + ops["DADD"])
    print("   ", ", ".join("%s:%d" % kv for kv in ops.most_common(14)))
