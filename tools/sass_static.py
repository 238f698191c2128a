#!/usr/bin/env python
"""Static SASS size per source function of one kernel (no profile needed):  tools/sass_static.py LIB.so KERNEL_MANGLED_NAME
Uses nvdisasm -g line info (innermost inlined frame) like tools/sass_attrib.py."""
import os, re, subprocess, sys, tempfile
from collections import Counter
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_attrib import func_table

lib, kern = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
cur, on, cnt, tabs = ("?", 0), False, Counter(), {}
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
        f, ln = cur
        if f not in tabs:
            tabs[f] = func_table(f)
        name = "?"
        for first, n in tabs[f]:
            if first <= ln:
                name = n
        cnt[(os.path.basename(f), name)] += 1
tot = sum(cnt.values())
print("total", tot)
for (f, n), c in cnt.most_common(40):
    print("%6d  %5.1f%%  %s:%s" % (c, 100.0 * c / tot, f, n))
