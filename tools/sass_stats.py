#!/usr/bin/env python
"""Dev tool: per-function SASS instruction mix of a built library (cuobjdump -sass).
usage: python tools/sass_stats.py [lib.so] [name-filter]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gym-solarpvder-environment_b200/csrc/libpvder_b200.so"
flt = sys.argv[2] if len(sys.argv) > 2 else "step_kernel"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None
stats = collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        stats[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        op = m.group(1).split(".")[0]
        stats[fn][op] += 1
for fn, c in stats.items():
    if flt not in fn:
        continue
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:90]
    tot = sum(c.values())
    f64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
    print(f"{dem}\n   total {tot}  fp64 {f64}  LDL {c['LDL']} STL {c['STL']}  LDS {c['LDS']} STS {c['STS']} MUFU {c['MUFU']} SEL/FSEL {c['SEL']+c['FSEL']} MOV {c['MOV']} IMAD {c['IMAD']} BRA {c['BRA']} CALL {c['CALL']}")
