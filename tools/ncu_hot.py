#!/usr/bin/env python
"""Hot source lines of a captured kernel: python tools/ncu_hot.py <file.ncu-rep> [top]  (needs -lineinfo; reads ncu --page source --csv)"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# the csv has per-kernel header blocks; find header line starting with "Address" or "Line"
lines = txt.splitlines()
hdr_i = [i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"') or l.startswith('"Line')]
print("header rows:", hdr_i[:3], file=sys.stderr)
rows = list(csv.reader(lines[hdr_i[0]:]))
hdr = rows[0]
print(hdr[:8], file=sys.stderr)
ci = {h: i for i, h in enumerate(hdr)}
samp = ci.get("# Samples") or ci.get("Warp Stall Sampling (All Samples)")
agg = collections.Counter(); inst = collections.Counter()
src_col = 1   # first "Source" column = CUDA source line (the second one is the SASS text)
for r in rows[1:]:
    if len(r) <= samp: continue
    try: s = float(r[samp] or 0)
    except ValueError: continue
    agg[(r[0] + ": " + r[src_col].strip())[:150]] += s
tot = sum(agg.values())
for k, v in agg.most_common(top): print("%6.2f%%  %s" % (100 * v / tot, k))
