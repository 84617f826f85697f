#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from the in-tree library:
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAPF (TMA load / store / L2 prefetch), UTCBAR (tcgen05.commit),
HMMA (legacy mma.sync / wmma).  Usage: python tools/sass_mnemonics.py [path/to/lib.so] > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UTCBAR|HMMA|HGMMA|LDGSTS)\b[.\w]*")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "diffpir_b200", "libdiffpir_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()     # noqa: E731
    counts = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn:
            m = PAT.search(line)
            if m:
                counts[fn][m.group(0)] += 1
    print(f"# SASS mnemonics per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; sm_100a)")
    print("# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG/UTMASTG/UTMAPF = TMA load/store/L2 prefetch, UTCBAR = tcgen05.commit, HMMA = mma.sync/wmma")
    for fn, c in counts.items():
        name = re.sub(r"CUtensorMap_st", "CUtensorMap", demangle(fn))
        name = re.sub(r"\(.*", "", name)
        body = ", ".join(f"{k} x{v}" for k, v in sorted(c.items())) if c else "-"
        print(f"{name:60s} {body}")


if __name__ == "__main__":
    main()
