"""Per-kernel SASS opcode histogram of libwwb200.so (the evidence that the hot kernels are tcgen05 / TMEM / bulk-copy
code): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,
SYNCS = mbarrier, MUFU / DFMA / FFMA2 the epilogue arithmetic.

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wakeword_detection_b200", "libwwb200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "MUFU", "FFMA2", "FADD2", "FFMA", "DFMA", "DADD", "DMUL",
        "HMMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kern, hist, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
            hist[kern] = collections.Counter()
            total[kern] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and kern:
            op = m.group(1)
            total[kern] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    hist[kern][k] += 1
                    break
    print("# SASS opcode histogram per kernel of wakeword_detection_b200/libwwb200.so (sm_100a), instruction counts in the binary")
    print("# %-44s %7s  %s" % ("kernel", "instrs", "  ".join(KEYS)))
    for k, h in hist.items():
        print("%-46s %7d  %s" % (k.replace("wwb::", "")[:46], total[k], "  ".join("%*d" % (len(n), h.get(n, 0)) for n in KEYS)))
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    print("%-46s %7d  %s" % ("TOTAL", sum(total.values()), "  ".join("%*d" % (len(n), tot.get(n, 0)) for n in KEYS)))


if __name__ == "__main__":
    main()
