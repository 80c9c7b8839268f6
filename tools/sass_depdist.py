"""Static dependency distance of the FP64 instructions of one kernel: for every DFMA / DMUL / DADD, how many instructions
earlier the nearest producer of one of its register sources was issued (1 = the instruction right before it).  A proxy
for the fixed-latency ("wait") stalls of a kernel that runs few warps per scheduler: the FP64 result latency is ~8 issue
slots, so distances below 4 stall a lone warp.
usage: python tools/sass_depdist.py <object-or-library> <kernel-name-substring>"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    if sys.argv[2] not in name:
        continue
    last, dist, n_ins = {}, [], 0
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", line)
        if not m:
            continue
        n_ins += 1
        op, ops = m.group(2), [o.strip() for o in m.group(3).split(",")]
        if op.split(".")[0] in ("DFMA", "DMUL", "DADD"):
            srcs = []
            for o in ops[1:]:
                mm = re.match(r"R(\d+)", o.lstrip("-|"))
                if mm:
                    srcs += ["R%d" % int(mm.group(1)), "R%d" % (int(mm.group(1)) + 1)]
            dist.append(min([n_ins - last[s] for s in srcs if s in last] or [99]))
        mm = re.match(r"R(\d+)", ops[0]) if ops and ops[0] else None
        if mm and not op.startswith(("ST", "BRA", "RED", "ATOM")):
            r = int(mm.group(1))
            wide = 4 if ".128" in op else 2 if (op[0] == "D" or ".64" in op) else 1
            for k in range(wide):
                last["R%d" % (r + k)] = n_ins
    c = collections.Counter(min(d, 9) for d in dist)
    tot = len(dist)
    print(name[:100])
    print("  instructions %d, FP64 %d; dependency distance histogram (9 = 9 or more):" % (n_ins, tot))
    print("  " + "  ".join("%d:%d (%.0f%%)" % (k, v, 100.0 * v / tot) for k, v in sorted(c.items())))
    print("  mean exposed slots per FP64 instruction if alone (latency 4 issue slots): %.2f" % (sum(max(0, 4 - d) for d in dist) / tot))
