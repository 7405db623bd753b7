"""Instruction histogram per kernel of libdensehead.so (cuobjdump -sass; runs without a GPU): the mnemonics that show
what the kernels are made of -- UBLKCP (1-D TMA bulk copies), FFMA2/FMUL2/FADD2 (Blackwell packed fp32), MUFU, REDUX,
shared-memory atomics, LDL/STL (register spills), UTMALDG/UTMASTG and UTC*MMA (tensor TMA / tcgen05: none expected, the
path has no contraction and its tiles are contiguous 1-D ranges).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cv-lite-object-detection_b200", "lib", "libdensehead.so")
WANT = ["UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "REDUX", "ATOMS", "ATOMG", "RED",
        "LDG", "STG", "LDS", "STS", "LDL", "STL", "SYNCS", "UCGABAR", "BAR", "VOTE", "SHFL", "DFMA", "DMUL", "DADD"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op = m.group(1)
            for w in WANT:
                if op == w:
                    cur[w] += 1
    names = demangle(list(kernels))
    print("libdensehead.so: cubins for %s; %d kernels" % (", ".join(arch), len(kernels)))
    cols = [w for w in WANT if any(k[w] for k in kernels.values())]
    print("%-110s %7s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    tot = collections.Counter()
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", names[name]).replace("dh::", "").replace("void ", "")
        print("%-110s %7d " % (short[:110], c["_total"]) + " ".join("%7d" % c[w] for w in cols))
        tot.update(c)
    print("%-110s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[w] for w in cols))
    absent = [w for w in ("UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA") if not tot[w]]
    print("absent (as expected: contiguous 1-D tiles, no contraction): " + ", ".join(absent))


if __name__ == "__main__":
    main()
