#!/usr/bin/env python
"""Instruction mix of the streaming loop of a particle-pass kernel, from the SASS of the built objects.

  python tools/sass_loop_count.py <object.o> <substring of the mangled kernel name> [marker ...]

The loop is the smallest backward-branch body that contains every marker (default: STG and DFMA.RM, i.e. the
particle loop that stores results and splits a cell coordinate).  One trip handles two particles per thread."""
import re
import subprocess
import sys


def loops(obj, name, markers):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out = []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        fname = f.split("\n", 1)[0]
        if name not in fname:
            continue
        ins = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip() for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        addr = lambda s: int(re.search(r"/\*([0-9a-f]{4,})\*/", s).group(1), 16)
        index = {addr(x): k for k, x in enumerate(ins)}
        best = None
        for i, l in enumerate(ins):
            m = re.search(r"\bBRA(?:\.[A-Z]+)* (0x[0-9a-f]+)", l)
            if not m or "BRA.U" in l:
                continue
            t = int(m.group(1), 16)
            if t < addr(l) and t in index:
                body = ins[index[t]:i + 1]
                if all(any(mk in x for x in body) for mk in markers) and (best is None or len(body) < len(best)):
                    best = body
        if best:
            out.append((fname, best))
    return out


def main():
    obj, name = sys.argv[1], sys.argv[2]
    markers = sys.argv[3:] or ["STG", "DFMA.RM"]
    for fname, body in loops(obj, name, markers):
        n = len(body)
        cnt = lambda pat: sum(1 for x in body if re.search(pat, x))
        f64 = cnt(r"\b(DFMA|DADD|DMUL)")
        print(fname)
        print(f"  loop: {n} instructions per trip = {n / 2:.1f} per particle; fp64 {f64 / 2:.1f}, LDS {cnt(r'LDS') / 2:.1f}, STS {cnt(r'STS') / 2:.1f}, "
              f"STG {cnt(r'STG') / 2:.1f}, LDG {cnt(r'LDG') / 2:.1f}, integer/move/control {(n - f64 - cnt(r'LDS|STS|STG|LDG')) / 2:.1f}")


if __name__ == "__main__":
    main()
