#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one captured kernel:
   python tools/ncu_lines.py <file.ncu-rep> [kernel-regex] [top]     (needs -lineinfo and ncu --import-source on)"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; kf = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["--kernel-name", "regex:" + kf, "--launch-count", "1"] if kf else ["--launch-count", "1"])
lines = subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()
h = [i for i, l in enumerate(lines) if l.startswith('"Line No"')][0]
rows = list(csv.reader(lines[h:])); hdr = rows[0]; ci = {}
for i, x in enumerate(hdr):
    ci.setdefault(x, i)
ie, sm = ci["Thread Instructions Executed"], ci["Warp Stall Sampling (All Samples)"]
agg, ags, cur = collections.Counter(), collections.Counter(), None
for r in rows[1:]:
    if len(r) <= ie or r[0] in ("File Path", "Function Name", "Line No"):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        continue
    if r[0] == "":
        continue                       # SASS row (already summed into its source line)
    key = "%s:%s  %s" % (cur, r[0], r[1].strip()[:120])
    try:
        a, b = float(r[ie] or 0), float(r[sm] or 0)
    except ValueError:
        continue                       # a source line whose quotes broke the CSV row (inline asm): its SASS is tiny
    agg[key] += a; ags[key] += b
tot, ts = sum(agg.values()) or 1, sum(ags.values()) or 1
print("thread instructions %.4e, stall samples %d" % (tot, ts))
for k, v in agg.most_common(top):
    print("%6.2f%% inst %6.2f%% stall  %s" % (100 * v / tot, 100 * ags[k] / ts, k))
