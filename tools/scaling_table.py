#!/usr/bin/env python
"""Markdown table of tools/scaling_study.sh's lines (profiles/r2_scaling.jsonl): per workload, frame ms / Grays/s / speed-up at every N measured."""
import collections, json, sys
rows = collections.OrderedDict()
for l in open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r2_scaling.jsonl"):
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    w = d["config"]["workload"].split(" (")[0].replace("3840x2160 16spp", "").replace(" photons cast per light", "").strip().replace(" ,", ",")
    rows.setdefault(w, {})[d["n_gpus"]] = d          # a later line of the same (workload, N) replaces an earlier one
ns = sorted({n for r in rows.values() for n in r})
print("| workload | " + " | ".join("N=%d ms / Grays/s%s" % (n, "" if n == 1 else " / speed-up") for n in ns) + " | frame CRC equal |")
print("|---|" + "---|" * (len(ns) + 1))
for w, r in rows.items():
    base = r.get(1)
    cells = []
    for n in ns:
        d = r.get(n)
        if d is None:
            cells.append("—")
        else:
            sp = "" if n == 1 or base is None else " / %.2f×" % (base["ms_per_step"] / d["ms_per_step"])
            cells.append("%.1f / %.2f%s" % (d["ms_per_step"], d["value"] / 1e3, sp))
    crc = {d.get("frame_crc32") for d in r.values()}
    print("| %s | %s | %s |" % (w, " | ".join(cells), "yes" if len(crc) == 1 else "NO " + str(crc)))
