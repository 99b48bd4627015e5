#!/usr/bin/env python
"""Dev tool: local-memory instructions (LDL/STL), calls and branches inside the biggest loop of one SASS function.
usage: python tools/sass_local.py lib.so name-filter"""
import re, subprocess, sys
lib, flt = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, body = None, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1); body[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m and fn:
        body[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in body.items():
    if flt not in fn: continue
    best = None
    for a, t in ins:
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m2 and int(m2.group(1), 16) < a:
            tgt = int(m2.group(1), 16)
            if best is None or a - tgt > best[1] - best[0]: best = (tgt, a)
    print(fn[:90], "loop %#x..%#x" % best)
    for a, t in ins:
        if best[0] <= a <= best[1] and re.search(r"LDL|STL|CALL|BSSY|BSYNC|BRA", t):
            print("  %#x %s" % (a, t))
