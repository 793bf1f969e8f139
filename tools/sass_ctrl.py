#!/usr/bin/env python
"""Decode the scoreboard fields of a kernel's SASS (cuobjdump -sass): stall count, write/read barrier, wait mask.
usage: python tools/sass_ctrl.py lib.so KERNEL_SUBSTRING [OPCODE_REGEX]   -- prints instructions whose wait mask is non-zero"""
import re, subprocess, sys
lib, kern = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on, ins, i = False, [], 0
while i < len(txt):
    if "Function :" in txt[i]:
        on = kern in txt[i]
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', txt[i]) if on else None
    if m and i + 1 < len(txt):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', txt[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((m.group(1), m.group(2), (hi >> 41) & 0xf, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
            i += 2
            continue
    i += 1
print(len(ins), "instructions")
for k, (ad, t, st, wb, rb, w) in enumerate(ins):
    if (pat and pat.search(t)) or (not pat and w):
        print(k, ad, t[:80].ljust(80), "stall", st, "wb", wb if wb != 7 else '-', "rb", rb if rb != 7 else '-', "wait", format(w, '06b'))
