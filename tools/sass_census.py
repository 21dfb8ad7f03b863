#!/usr/bin/env python
"""Per-kernel SASS census of libpcad.so: how many tensor-core (UTCHMMA / HMMA), TMEM (LDTM / STTM), TMA (UTMALDG / UTMASTG)
and special-function (MUFU.*) instructions each kernel contains -- the static proof that the GEMMs / the SSD scan run on
tcgen05 with TMEM accumulators and TMA staging, and where the MUFU work sits.

    python tools/sass_census.py [path/to/libpcad.so] > profiles/r02_sass_census.txt
"""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = ["UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP",
            "MUFU.RSQ", "MUFU.TANH", "FFMA2", "FMUL2", "FADD2", "LDGSTS", "BAR.SYNC"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "plantcaduceus_b200", "libpcad.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    sizes = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            sizes[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        sizes[cur] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + ".") or (p.startswith("MUFU") and op.startswith(p)):
                counts[cur][p] += 1
    names = demangle(list(counts))
    short = lambda n: re.sub(r"\(.*", "", names[n].replace("(anonymous namespace)::", "")).replace("pcad::", "").replace("void ", "")
    cols = [p for p in PATTERNS if any(c[p] for c in counts.values())]
    print(f"SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); instruction counts per kernel")
    print(f"{'kernel':<64} {'instrs':>7} " + " ".join(f"{c:>8}" for c in cols))
    total = collections.Counter()
    for n, c in counts.items():
        if sizes[n] == 0:
            continue
        print(f"{short(n)[:64]:<64} {sizes[n]:>7} " + " ".join(f"{c[p]:>8}" for p in cols))
        total.update(c)
    print(f"{'TOTAL':<64} {sum(sizes.values()):>7} " + " ".join(f"{total[p]:>8}" for p in cols))


if __name__ == "__main__":
    main()
