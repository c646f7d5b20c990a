#!/usr/bin/env python
"""Top source lines of an ncu report by stall samples (needs -lineinfo). usage: ncu_by_line.py <rep> [n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = collections.OrderedDict()
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 7 and r[0].strip().isdigit():
        try: s, k = int(r[6]), int(r[7])
        except ValueError: continue
        key = (cur, int(r[0]))
        e = agg.setdefault(key, [0, 0, r[1]])
        e[0] += s; e[1] += k
tot = sum(v[0] for v in agg.values()) or 1
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print("%5.1f%% %10d  %s:%d  %s" % (100.0 * v[0] / tot, v[1], f, l, v[2].strip()[:120]))
