#!/usr/bin/env python3
"""The comparison rule of the reference's `dnadiff` (dynadjust/dnadiff/dnadiff.cpp:40-268), restated: two DynAdjust
output files match when, line by line after the skipped part, they have the same tokens — numeric tokens within a
tolerance, everything else exactly; lines that carry times, versions, paths, thread counts or the largest-correction
message are not compared.  The reference's CI applies it to dnaadjust's .adj files (CMakeLists.txt:1033, 1190-1191).

    tools/dnadiff.py ours.adj expected.adj [--skip-headers N] [--skip-to-marker TEXT] [-t TOL] [-v]
"""
import argparse
import re
import sys

NUMERIC = re.compile(r"^[-+]?[0-9]*\.?[0-9]+([eE][-+]?[0-9]+)?$")
SKIP = ("File created:", "Build:", "Version:", "time", "File name:", "Input files:", "Output folder:", "Input folder:", "Command line arguments:",
        "threads", "Maximum station correction", "(e, n, up)")


def should_skip(line):
    return bool(line) and any(k in line for k in SKIP)


def compare(file1, file2, tolerance=0.001, skip_headers=0, skip_to_marker="", verbose=False, out=sys.stdout):
    def lines(path):
        with open(path, errors="replace") as f:
            return f.read().split("\n")
    a, b = lines(file1), lines(file2)
    if a and a[-1] == "":
        a.pop()
    if b and b[-1] == "":
        b.pop()
    if skip_to_marker:
        for name, seq in ((file1, a), (file2, b)):
            k = next((i for i, l in enumerate(seq) if skip_to_marker in l), None)
            if k is None:
                print(f'Error: Marker "{skip_to_marker}" not found in {name}', file=out)
                return 1 << 30
            del seq[:k + 1]
    a, b = a[skip_headers:], b[skip_headers:]
    differences = 0
    for n, (l1, l2) in enumerate(zip(a, b), start=skip_headers + 1):
        if (should_skip(l1) and should_skip(l2)) or (not l1 and not l2):
            continue
        t1, t2 = l1.split(), l2.split()
        ok = len(t1) == len(t2)
        for x, y in zip(t1, t2):
            if NUMERIC.match(x) and NUMERIC.match(y):
                if abs(float(x) - float(y)) > tolerance:
                    ok = False
                    break
            elif x != y:
                ok = False
                break
        if not ok:
            differences += 1
            if verbose:
                print(f"Line {n}:\n  {file1}: {l1}\n  {file2}: {l2}", file=out)
    if len(a) != len(b):
        differences += 1
        if verbose:
            print(f"Files have different number of lines ({len(a)} vs {len(b)})", file=out)
    return differences


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("file1")
    ap.add_argument("file2")
    ap.add_argument("-t", "--tolerance", "--tol", type=float, default=0.001)
    ap.add_argument("--skip-headers", type=int, default=0)
    ap.add_argument("--skip-to-marker", default="")
    ap.add_argument("-v", "--verbose", action="store_true")
    o = ap.parse_args()
    d = compare(o.file1, o.file2, o.tolerance, o.skip_headers, o.skip_to_marker, o.verbose)
    print(f"\nTolerance used: {o.tolerance}\n" + ("Files match within tolerance." if d == 0 else f"Files differ beyond tolerance ({d} lines)."))
    return 0 if d == 0 else 1


if __name__ == "__main__":
    try:
        sys.exit(main())
    except BrokenPipeError:      # output cut short by the reader (| head)
        sys.exit(1)
