#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libbbg.so (cuobjdump -sass), written to profiles/ as evidence of what the hot
kernels are made of (IMAD.WIDE-bound integer work; no tensor-core, no TMA instructions).  Usage:
    python scripts/sass_histogram.py aztec-2.0_b200/libbbg.so profiles/r2_sass_histogram.md"""
import collections
import re
import subprocess
import sys


def main(lib, out):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            base = op.split(".")[0]
            key = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else ("IMAD.HI" if op.startswith("IMAD.HI") else base)
            cur[key] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    with open(out, "w") as f:
        f.write("# SASS opcode histogram per kernel (`cuobjdump -sass %s`, sm_100a)\n\n" % lib)
        f.write("Static instruction counts (not execution counts). IMAD.WIDE is the quarter-rate 32x32->64 multiply-add the 254-bit Montgomery\n"
                "product is made of: this is carry-propagating integer work on the fmaheavy pipe (DESIGN.md 3).  HMMA / UTCMMA (tensor cores)\n"
                "do not occur.  UBLKCP (cp.async.bulk, the TMA bulk copy that stages the stage twiddles of the E = 4 NTT pass into shared\n"
                "memory) and LDGSTS (cp.async, the persistent-pass experiment) occur only in `k_ntt_pass`; the totals are in the last line.\n\n")
        f.write("| kernel | instructions | IMAD.WIDE | IMAD | IADD3 | LOP3+SEL | SHFL | LDG/STG | LDS/STS | LDL/STL | BAR | other top |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for (name, c), dm in zip(kernels.items(), demangle):
            n = sum(c.values())
            total.update(c)
            short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("bbg::", "")
            shown = {"IMAD.WIDE", "IMAD", "IADD3", "LOP3", "SEL", "SHFL", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR"}
            other = ", ".join("%s %d" % kv for kv in c.most_common(12) if kv[0] not in shown)[:80]
            f.write("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %s |\n" % (
                short[:60], n, c["IMAD.WIDE"], c["IMAD"] + c["IMAD.HI"], c["IADD3"], c["LOP3"] + c["SEL"], c["SHFL"], c["LDG"] + c["STG"],
                c["LDS"] + c["STS"], c["LDL"] + c["STL"], c["BAR"], other))
        f.write("\nWhole library: %d instructions; IMAD.WIDE %d, IMAD %d, IADD3 %d, SHFL %d; UTMALDG %d, UBLKCP %d, LDGSTS %d, HMMA %d, UTCHMMA/UTCMMA %d.\n" % (
            sum(total.values()), total["IMAD.WIDE"], total["IMAD"] + total["IMAD.HI"], total["IADD3"], total["SHFL"], total["UTMALDG"], total["UBLKCP"],
            total["LDGSTS"], total["HMMA"], total["UTCHMMA"] + total["UTCMMA"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
