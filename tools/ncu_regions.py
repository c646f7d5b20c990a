#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu source-page CSV: splits the kernel into regions by execution count and
prints sample / instruction shares, opcode mix and stall reasons per region.
usage: ncu_regions.py <rep> [hot_threshold_fraction]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; ins = []
for r in rows:
    if r and r[0] in ("Address", "Line No") or (len(r) > 3 and "Source" in r[:3] and "# Samples" in r):
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            ins.append(d)
        except Exception:
            pass
print("columns:", hdr[:12] if hdr else None, "n instr:", len(ins))
def gi(d, k):
    try: return int(d.get(k, "0") or 0)
    except ValueError: return 0
tot_s = sum(gi(d, "# Samples") for d in ins) or 1
tot_i = sum(gi(d, "Instructions Executed") for d in ins) or 1
mx = max(gi(d, "Instructions Executed") for d in ins)
print("total samples %d, warp instr %d, max exec count %d" % (tot_s, tot_i, mx))
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
# regions by exec count buckets
buckets = collections.OrderedDict()
def bucket(n):
    if n == 0: return "never"
    f = n / float(mx)
    if f > 0.5: return "hot(>50% of max)"
    if f > 0.1: return "warm(10-50%)"
    if f > 0.01: return "cool(1-10%)"
    return "cold(<1%)"
for d in ins:
    b = bucket(gi(d, "Instructions Executed"))
    e = buckets.setdefault(b, dict(n=0, s=0, i=0, ops=collections.Counter(), st=collections.Counter()))
    e["n"] += 1; e["s"] += gi(d, "# Samples"); e["i"] += gi(d, "Instructions Executed")
    op = (d.get("Source") or "").strip().split()
    op = [o for o in op if not o.startswith("@")]
    e["ops"][op[0].split(".")[0] if op else "?"] += gi(d, "Instructions Executed")
    for c in stall_cols: e["st"][c] += gi(d, c)
for b, e in buckets.items():
    print("\n== %s: %d static instr, %.1f%% of samples, %.1f%% of executed warp-instr" % (b, e["n"], 100.0 * e["s"] / tot_s, 100.0 * e["i"] / tot_i))
    print("   opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(e["i"], 1)) for k, v in e["ops"].most_common(14)))
    ts = sum(e["st"].values()) or 1
    print("   stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / ts) for k, v in e["st"].most_common(8)))
