#!/usr/bin/env python
"""Per-kernel resource table (registers, local frame, spills, shared memory) of libdrt.so from `nvcc -Xptxas -v`.
usage: python tools/ptxas_report.py [-DNAME=VALUE ...]   (builds to a scratch file, never touches the in-tree library)"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_old_b200 import build as B


def report(defines=()):
    out = os.path.join(tempfile.mkdtemp(), "libdrt_report.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + B.NVCC_FLAGS + ["-D" + d for d in defines] + ["-Xptxas", "-v"] + [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-o", out] + B.LINK_FLAGS
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr); raise SystemExit(1)
    rows, cur = [], None
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"name": subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void drt::", ""), "stack": 0, "spill_st": 0, "spill_ld": 0, "regs": 0, "smem": 0}
            rows.append(cur); continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and not cur.get("seen"):       # the first properties block after the entry line is the kernel's own (callee blocks follow)
            cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups()); cur["seen"] = True
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1)); s = re.search(r"(\d+) bytes smem", line); cur["smem"] = int(s.group(1)) if s else 0
    return rows


if __name__ == "__main__":
    rows = report([a[2:] for a in sys.argv[1:] if a.startswith("-D")])
    print("| kernel | registers | local frame B | spill st/ld B | smem B |\n|---|---|---|---|---|")
    for r in sorted(rows, key=lambda r: r["name"]):
        print("| `%s` | %d | %d | %d / %d | %d |" % (r["name"], r["regs"], r["stack"], r["spill_st"], r["spill_ld"], r["smem"]))
