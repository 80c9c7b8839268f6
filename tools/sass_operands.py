"""FP64 operand mix of a kernel from SASS: DFMA / DMUL / DADD instructions by the number of distinct vector-register
source operands (R#; uniform registers UR#, constants c[..] and immediates come through the uniform datapath).
tools/fp64_peak.cu measures the issue rate of each class on B200: a warp-wide FP64 instruction occupies the pipe for
2.0 cycles with one vector-register source, 2.58 with two and 3.75 with three (34.2 / 26.5 / 18.2 TFLOP/s DFMA), so the
pipe-bound time of a kernel is  Σ_class count × cycles, not count × 2.
usage: python tools/sass_operands.py [lib.so] [kernel substring] [cycles1,cycles2,cycles3]"""
import collections
import json
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "ilqr.jl_b200/libilqr_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "round_lpt_two_linkILi12ELi4ELb0"
cyc = [float(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2.0, 2.58, 3.75]


def classify(so, pat):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        if pat not in name:
            continue
        cnt = collections.Counter()
        for line in f.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?(DFMA|DMUL|DADD)(?:\.[A-Z0-9.]+)?\s+([^;]+);", line)
            if not m:
                continue
            ops = [o.strip() for o in m.group(2).split(",")][1:]          # sources
            regs = set()
            for o in ops:
                o = o.lstrip("-|").rstrip("|")
                mm = re.match(r"R(\d+)", o)
                if mm:
                    regs.add(mm.group(1))
            cnt[(m.group(1), max(1, len(regs)))] += 1
        out[name] = cnt
    return out


if __name__ == "__main__":
    for name, cnt in classify(so, pat).items():
        by = collections.Counter()
        for (op, k), v in cnt.items():
            by[k] += v
        tot = sum(by.values())
        cycles = sum(by[k] * cyc[k - 1] for k in by)
        print(json.dumps({"kernel": name[-70:], "fp64": tot, "by_vector_register_sources": {str(k): by[k] for k in sorted(by)},
                          "by_op": {"%s/%d" % k: v for k, v in sorted(cnt.items())},
                          "pipe_cycles_per_static_pass": cycles, "mean_cycles_per_fp64_instr": cycles / tot,
                          "effective_peak_fraction_of_uniform_operand_peak": 2.0 * tot / cycles}))
