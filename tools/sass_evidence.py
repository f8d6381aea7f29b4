#!/usr/bin/env python
"""Count the SASS mnemonics that show which hardware paths the built kernels use (profiles/r1_sass_evidence.txt):
UBLKCP = cp.async.bulk through the TMA engine, SYNCS.* = mbarrier operations, ACQBULK / PREEXIT = programmatic
dependent launch (griddepcontrol.wait / launch_dependents), DFMA.RZ / DFMA.RM = the one-instruction floor of the
cell locate, ATOMS.CAST.SPIN = CAS-loop fp64 shared-memory atomics (histogram fallback paths only).
Usage: python tools/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "vlasovparticlemethods.jl_b200", "csrc", "build")
PAT = re.compile(r"\b(UBLKCP[.\w]*|SYNCS[.\w]*|ACQBULK[.\w]*|PREEXIT[.\w]*|DFMA\.R[ZMP][.\w]*|ATOMS\.CAST[.\w]*|LDG\.E[.\w]*128[.\w]*|STG\.E[.\w]*128[.\w]*)")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"vpm::\(anonymous namespace\)::", "", o).split("(")[0].replace("void ", "") for o in out]


def main():
    print(__doc__.strip().split("Usage")[0].strip())
    for obj in ("kernels_vp.o", "kernels_lb.o"):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
        counts = collections.OrderedDict()
        fn = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = m.group(1)
                counts[fn] = collections.Counter()
                continue
            m = PAT.search(line)
            if m and fn:
                counts[fn][m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("DFMA", "ATOMS", "SYNCS")) else "")] += 1
        names = list(counts)
        print(f"\n## {obj}")
        for raw, nice in zip(names, demangle(names)):
            if not re.search(r"<4,|field", nice):      # the order-4 instantiations and the field kernels
                continue
            c = counts[raw]
            if c:
                print(f"{nice:58s} " + "  ".join(f"{k} {v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
