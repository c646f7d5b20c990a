#!/usr/bin/env python
"""Aggregate an ncu source page (CUDA-source view, needs -lineinfo) by enclosing function of a given .cuh/.cu file.
usage: ncu_by_function.py <rep> <source-file>"""
import csv, io, re, subprocess, sys, collections
rep, srcfile = sys.argv[1], sys.argv[2]
lines = open(srcfile).read().split("\n")
# function start lines: "__device__ ... name(" or "__global__"
starts = []
for i, l in enumerate(lines, 1):
    m = re.match(r"^(template.*)?\s*__(device|global)__.*?\b([A-Za-z_0-9]+)\s*\(", l)
    if m and not l.startswith(" "):
        starts.append((i, m.group(3)))
def func_of(n):
    name = "?"
    for s, nm in starts:
        if s <= n: name = nm
        else: break
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0])
base = srcfile.split("/")[-1]
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 7 and r[0].strip().isdigit():
        try: s, n = int(r[6]), int(r[7])
        except ValueError: continue
        key = func_of(int(r[0])) if cur == base else "[" + str(cur) + "]"
        agg[key][0] += s; agg[key][1] += n
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %5.1f%% samples  %5.1f%% warp-instr  (%d instr)" % (k, 100.0 * v[0] / ts, 100.0 * v[1] / ti, v[1]))
