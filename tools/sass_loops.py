#!/usr/bin/env python3
"""List the loops (backward branches) of one kernel's SASS with their size and opcode mix.
usage: cuobjdump -sass -fun <mangled> lib.so > k.sass ; python tools/sass_loops.py k.sass [min_len]"""
import re, sys, collections
lines = open(sys.argv[1]).read().split('\n')
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ins = []
for l in lines:
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
def op(t):
    p = t.split()
    o = p[1] if p[0].startswith('@') else p[0]
    return o.split('.')[0]
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)', t)
    if not m: continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr2i:
        s = addr2i[tgt]
        n = i - s + 1
        if n < minlen: continue
        c = collections.Counter(op(x[1]) for x in ins[s:i + 1])
        print(f"loop {tgt:#x}..{a:#x}  {n} instr ({n*16} B): " + ' '.join(f"{k}:{v}" for k, v in c.most_common(14)))
