#!/usr/bin/env python
"""Dev tool: backward branches (loops) of one SASS function with the instruction mix of each loop body.
usage: python tools/sass_loops.py lib.so name-filter"""
import collections, re, subprocess, sys
lib, flt = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, body = None, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1); body[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m and fn:
        body[fn].append((int(m.group(1), 16), m.group(2)))
for fn, ins in body.items():
    if flt not in fn: continue
    print(subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:100])
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?", t)
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a:
                c = collections.Counter()
                for b, u in ins:
                    if tgt <= b <= a:
                        op = re.sub(r"^@!?U?P\d+\s+", "", u).split()[0].split(".")[0]
                        c[op] += 1
                f64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
                print(f"  loop {tgt:#x}..{a:#x}: {sum(c.values())} instr, fp64 {f64} (DFMA {c['DFMA']} DADD {c['DADD']} DMUL {c['DMUL']} DSETP {c['DSETP']}), LDL {c['LDL']} STL {c['STL']} LDC {c['LDC']+c['ULDC']} MUFU {c['MUFU']} FSEL/SEL {c['FSEL']+c['SEL']} MOV {c['MOV']+c['IMAD']} BRA {c['BRA']} CALL {c['CALL']}")
