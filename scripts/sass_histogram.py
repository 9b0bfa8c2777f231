#!/usr/bin/env python
"""SASS listing + opcode histogram of one kernel of a built object (evidence for profiles/).

    python scripts/sass_histogram.py OBJ MANGLED_NAME [--range 0xe60:0x21a0 ...] [--out FILE]

Prints the histogram of the whole kernel and, for each --range (byte addresses, end exclusive; several ranges are
added up), of that address range -- e.g. the tile loop.  Packed fp32x2 instructions (FFMA2 / FMUL2 / FADD2) occupy the
FMA pipe for two cycles per warp, every other FMA- or ALU-pipe instruction for one (scripts/ubench/), so the histogram
also reports "pipe units" = packed x 2 + the rest.
"""
import argparse
import collections
import re
import subprocess
import sys


def listing(obj, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
    rows = []
    for ln in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    return rows


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


def hist(rows):
    h = collections.Counter(opcode(t) for _, t in rows)
    packed = sum(h[k] for k in ("FFMA2", "FMUL2", "FADD2"))
    n = sum(h.values())
    return h, n, packed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("fun")
    ap.add_argument("--range", action="append", default=[])
    ap.add_argument("--label", default="")
    ap.add_argument("--listing", action="store_true", help="also print the full listing")
    a = ap.parse_args()
    rows = listing(a.obj, a.fun)
    print(f"# {a.label or a.fun}")
    h, n, packed = hist(rows)
    print(f"whole kernel: {n} instructions, {packed} packed fp32x2")
    if a.range:
        sel = []
        for r in a.range:
            lo, hi = (int(x, 16) for x in r.split(":"))
            sel += [(ad, t) for ad, t in rows if lo <= ad < hi]
        h, n, packed = hist(sel)
        print(f"ranges {' '.join(a.range)}: {n} instructions per warp-tile (8 pixels per lane), {packed} packed fp32x2, "
              f"pipe units = {n + packed}")
    for k, v in h.most_common():
        print(f"  {v:5d} {k}")
    if a.listing:
        print()
        for ad, t in rows:
            print(f"/*{ad:04x}*/ {t} ;")


if __name__ == "__main__":
    sys.exit(main())
