#!/usr/bin/env python
"""Static SASS evidence per kernel of the in-tree library: instruction counts of the mnemonics that matter (tensor-memory
loads / stores, TMA bulk / tensor copies, fp64 arithmetic, 256-bit stores), registers, code size.  CPU only (cuobjdump)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "autonomous-racing-lpv-mpp-mpc_b200", "liblpvmpc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for l in res.split("\n"):
    m = re.search(r"Function (\S+):", l)
    if m: cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", l)
    if m and cur: regs[cur] = tuple(int(x) for x in m.groups())
want = ["LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMASTG", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "STG.E.ENL2.256", "SHFL", "LDS", "STS", "BAR.SYNC", "WARPSYNC", "NANOSLEEP"]
cnt = collections.OrderedDict()
cur = None
for l in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        cnt[cur]["_n"] += 1
        for w in want:
            if re.search(r"\b" + re.escape(w), l): cnt[cur][w] += 1
print("arch:", re.search(r"arch = (\S+)", sass).group(1))
for f, c in cnt.items():
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
    r = regs.get(f, (0, 0, 0))
    print("\n%s\n  %d instructions (%.0f KB), %d registers, %d B stack, %d B static shared" % (name, c["_n"], c["_n"] * 16 / 1024, r[0], r[1], r[2]))
    print("  " + "  ".join("%s %d" % (w, c[w]) for w in want if c[w]))
