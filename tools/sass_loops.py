#!/usr/bin/env python
"""Per-loop instruction mix of a kernel's SASS (backward branches delimit loops): DMMA / LDS / LDGSTS / spills.
Usage: python tools/sass_loops.py build/csrc/kernels_gemm.o bsc_gemm_kernel"""
import re, subprocess, sys
obj, name = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins = []
on = False
for ln in out.split("\n"):
    if "Function :" in ln:
        on = name in ln
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:`\(\S+\)|0x([0-9a-f]+))", t)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            loops.append((addr[tgt], i))
print(f"{len(ins)} instructions, {len(loops)} loops")
for s, e in sorted(loops):
    body = [t for _, t in ins[s:e + 1]]
    c = lambda k: sum(k in t for t in body)
    if c("DMMA") or c("LDGSTS") or c("STL") or c("LDL"):
        print(f"  loop [{s:6d},{e:6d}] len {e-s+1:5d}: DMMA {c('DMMA'):3d} LDS {c('LDS'):3d} LDGSTS {c('LDGSTS'):3d} STL {c('STL'):3d} LDL {c('LDL'):3d} SYNCS {c('SYNCS'):2d} BAR {c('BAR.'):2d}")
