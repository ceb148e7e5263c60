"""Per-kernel SASS opcode census of the built library (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that prove
which kernels run on the Blackwell tensor pipe (UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier ops) next to legacy HMMA (mma.sync) and MUFU counts.

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.tsv
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffews_b200", "libdiffews_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU", "FFMA2", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["total"] += 1
            for o in OPS[:-1]:
                if op.startswith(o):
                    counts[kern][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass diffews_b200/libdiffews_b200.so (sm_100a): instruction counts per kernel")
    print("kernel\t" + "\t".join(OPS))
    for (k, c), name in zip(counts.items(), demangle):
        name = re.sub(r"\(anonymous namespace\)::|dfw::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        print(name + "\t" + "\t".join(str(c[o]) for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
