#!/usr/bin/env python
"""Summarise the source page of an .ncu-rep: instruction counts per unit of work, stall reasons, hot spots.
usage: python tools/ncu_src.py rep.ncu-rep UNITS [min_pct]"""
import collections, csv, io, subprocess, sys
rep, T = sys.argv[1], float(sys.argv[2])
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
hi = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
rd = csv.reader(io.StringIO("\n".join(lines[hi:])))
hdr = next(rd)
rows = [r for r in rd if len(r) == len(hdr)]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in rows)
print("instructions", len(rows), "executed/unit %.1f" % (sum(int(r[iexec]) for r in rows) / T), "samples", tot)
st = collections.Counter()
for r in rows:
    for i, h in stall_cols:
        st[h] += int(r[i] or 0)
print("stalls:", ", ".join("%s %.1f%%" % (h, 100 * c / max(1, sum(st.values()))) for h, c in st.most_common(9)))
oc, ocs = collections.Counter(), collections.Counter()
for r in rows:
    op = [x for x in r[isrc].strip().split() if not x.startswith('@')][0].split('.')[0]
    oc[op] += int(r[iexec]); ocs[op] += int(r[isamp])
print("opcodes:", ", ".join("%s %.1f/u %.1f%%" % (op, c / T, 100 * ocs[op] / tot) for op, c in oc.most_common(16)))
for k, r in enumerate(rows):
    s = int(r[isamp])
    if s > tot * minpct / 100:
        top = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        print("%5d %-64s %5.1f%% exec/u %.2f  %s" % (k, r[isrc].strip()[:64], 100 * s / tot, int(r[iexec]) / T,
                                                    " ".join("%s:%d" % (h[6:], c) for c, h in top)))
