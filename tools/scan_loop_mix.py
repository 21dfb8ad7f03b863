#!/usr/bin/env python
"""Static instruction mix of the selective scan's main loop, from the SASS of the built library (no GPU needed).

biscan_kernel's recurrence runs in 8-step blocks (scan.cuh); this finds those loops in `cuobjdump -sass` (the innermost
backward branches whose body holds 8 x 17 MUFU.EX2) and prints, per step and warp, how many issue slots the loop needs
against how many cycles its MUFU work occupies the quarter-SM special-function pipe (16 lanes / clk / SM = 4 per scheduler:
a warp-wide MUFU instruction holds it for 8 cycles).  That ratio is the issue-slot utilisation a MUFU-saturated loop shows
in ncu (`smsp__issue_active`), and what is left of the pipe's time is chunk prologue / epilogue / barrier time.

    python tools/scan_loop_mix.py > profiles/r02_scan_loop_mix.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "plantcaduceus_b200", "libpcad.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = {}
    print(f"main-loop instruction mix of biscan_kernel in {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        if "biscan_kernel" not in name:
            continue
        try:
            names[name] = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        except OSError:
            names[name] = name
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
        loops = []
        for a, t in ins:
            m = re.search(r"\bBRA\S*\s+(?:[!\w]+,\s*)?`?\(?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
        print(f"\n{names[name].replace('pcad::', '').replace('void ', '')}   ({len(ins)} instructions)")
        for tgt, a in loops:
            body = [t.split()[0] for x, t in ins if tgt <= x <= a]
            ops = collections.Counter(op if op.startswith("MUFU") else op.split(".")[0] for op in body)
            mufu = sum(v for k, v in ops.items() if k.startswith("MUFU"))
            if ops.get("MUFU.EX2", 0) != 8 * 17 or len(body) > 1000:
                continue                                            # not an 8-step block of the recurrence
            per_step, mufu_step = len(body) / 8.0, mufu / 8.0
            print(f"  8-step block {tgt:#x}..{a:#x}: {len(body)} instructions = {per_step:.1f} per step and warp")
            print("    " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
            print(f"    MUFU per step {mufu_step:.0f} (16 decays 2^(d*A_n) + 1 softplus) -> {8 * mufu_step:.0f} pipe cycles per step and warp; "
                  f"issue slots per step {per_step:.1f}")
            print(f"    MUFU-saturated loop => issue-slot utilisation {per_step / (8 * mufu_step):.2f} "
                  f"(ncu smsp__issue_active of the whole kernel: 0.58), issue headroom {1 - per_step / (8 * mufu_step):.2f}")


if __name__ == "__main__":
    main()
