#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of an ncu capture to source functions.

    python tools/sass_attrib.py REPORT.ncu-rep LIB.so KERNEL_MANGLED_NAME [--lines]

ncu's source page (CSV) gives, per SASS instruction of the profiled kernel, executed-instruction counts and
warp-stall samples; nvdisasm -g gives the source line of every SASS instruction of the same cubin.  The two
are joined by instruction index and summed per source function (the innermost inlined frame)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def func_table(path):
    """[(first_line, name)] of function definitions in a source file (heuristic: '__device__' / '__global__' lines)."""
    tab = []
    try:
        src = open(path, encoding="utf-8", errors="replace").read().split("\n")
    except OSError:
        return tab
    pat = re.compile(r"(?:__device__|__global__)[^;(]*?([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    for n, line in enumerate(src, 1):
        if line.lstrip().startswith("//"):
            continue
        m = pat.search(line)
        if m:
            tab.append((n, m.group(1)))
    return tab


def main():
    rep, lib, kern = sys.argv[1:4]
    by_line = "--lines" in sys.argv
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    lines, cur, on = [], ("?", 0), False
    for l in dis:
        if l.startswith(".text."):
            on = l.strip() == ".text.%s:" % kern
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    col = {k: hdr.index(k) for k in ("Instructions Executed", "# Samples", "stall_no_inst", "stall_wait", "stall_short_sb",
                                       "stall_long_sb", "stall_branch_resolving", "Source")}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    if len(data) != len(lines):
        print("warning: %d SASS rows in the report vs %d in the cubin (different build?)" % (len(data), len(lines)), file=sys.stderr)
    tabs = {}
    agg = defaultdict(lambda: [0] * 8)
    for r, (f, ln) in zip(data, lines):
        if f not in tabs:
            tabs[f] = func_table(f)
        name = "?"
        for s, nme in tabs[f]:
            if s <= ln:
                name = nme
            else:
                break
        key = "%s:%s" % (os.path.basename(f), ("%d" % ln) if by_line else name)
        a = agg[key]
        a[0] += 1
        for j, k in enumerate(("Instructions Executed", "# Samples", "stall_no_inst", "stall_wait", "stall_short_sb",
                               "stall_long_sb", "stall_branch_resolving"), 1):
            a[j] += int(r[col[k]] or 0)
    tot = [sum(a[j] for a in agg.values()) for j in range(8)]
    print("| function | SASS instrs | executed | exec %% | samples %% | no_inst %% | wait %% | short_sb %% | long_sb %% | branch %% |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if a[1] == 0 and a[2] == 0:
            continue
        print("| %s | %d | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
            k, a[0], a[1], 100 * a[1] / max(tot[1], 1), 100 * a[2] / max(tot[2], 1), 100 * a[3] / max(tot[3], 1),
            100 * a[4] / max(tot[4], 1), 100 * a[5] / max(tot[5], 1), 100 * a[6] / max(tot[6], 1), 100 * a[7] / max(tot[7], 1)))
    print("| TOTAL | %d | %d | | %d | %d | %d | %d | %d | %d |" % tuple(tot))


if __name__ == "__main__":
    main()
