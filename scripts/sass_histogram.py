"""Opcode histogram of the Blackwell-specific SASS in libflowcon_b200.so, per kernel (evidence that the hot kernels are
tcgen05 / TMEM / TMA code: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA).
    python scripts/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flowconductor_b200", "lib", "libflowcon_b200.so")
INTERESTING = re.compile(r"^(UTC\w*MMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|SYNCS|F2FP|FFMA2|FMUL2|FADD2|HMMA|"
                         r"MUFU|USETMAXREG|UCGABAR_ARV|ELECT|REDUX|LDS|STS|LDG|STG|FFMA|FMNMX3?)$")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn = None
    hist = collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            if INTERESTING.match(op):
                hist[fn][op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(hist.keys()), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass flowconductor_b200/lib/libflowcon_b200.so — Blackwell-specific and hot opcodes per kernel")
    total = collections.Counter()
    for mangled, name in sorted(zip(hist.keys(), demangle), key=lambda t: t[1]):
        c = hist[mangled]
        total.update(c)
        if not any(k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLK")) for k in c):
            continue
        print(name[:150])
        print("    " + ", ".join("{} x{}".format(k, v) for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
    print("# whole library")
    print("    " + ", ".join("{} x{}".format(k, v) for k, v in sorted(total.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    sys.exit(main())
