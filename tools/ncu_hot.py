#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples by CUDA source line.

ncu's CSV source page is per SASS instruction; line info comes from `nvdisasm -g` on the cubin extracted from the
shared library (compile with -lineinfo).  usage: tools/ncu_hot.py report.ncu-rep lib.so kernel_substring [top_n]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_samples(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        st = {h[6:]: int(r[i]) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i])}
        res.append((int(r[0], 16), r[col["Source"]].strip(), int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0), st))
    return res


def line_table(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    table = {}
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        cur, line, fn = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                fn, line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
            if m and cur and kernel_sub in cur:
                table[int(m.group(1), 16)] = (fn, line)
    return table


def main():
    rep, lib, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sass = sass_samples(rep)
    base = sass[0][0]
    table = line_table(lib, ksub)
    agg = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot = 0
    for addr, text, n, ie, st in sass:
        key = table.get(addr - base, ("?", 0))
        a = agg[key]
        a[0] += n
        a[1] += ie
        for k, v in st.items():
            a[2][k] += v
        tot += n
    srcs = {}
    print("total samples %d, %d SASS instructions, %d mapped" % (tot, len(sass), sum(1 for a, *_ in sass if (a - base) in table)))
    for (fn, line), (n, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if fn not in srcs:
            path = os.path.join(os.path.dirname(os.path.abspath(lib)), fn)
            srcs[fn] = open(path).read().splitlines() if os.path.exists(path) else []
        text = srcs[fn][line - 1].strip()[:90] if 0 < line <= len(srcs[fn]) else ""
        top3 = " ".join("%s=%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%6d %5.1f%% inst=%-9d %s:%-4d %-90s %s" % (n, 100.0 * n / max(tot, 1), ie, fn, line, text, top3))


if __name__ == "__main__":
    main()
