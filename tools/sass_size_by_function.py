#!/usr/bin/env python
"""Static SASS size of one kernel by source function: nvdisasm -g -c output (line markers), instructions attributed to the
enclosing function of the marked source line.  usage: sass_size_by_function.py <all.sass> <kernel-substring>"""
import re, sys, collections, os
sass, kern = sys.argv[1], sys.argv[2]
srcs = {}
def func_of(path, n):
    if path not in srcs:
        starts = []
        try:
            for i, l in enumerate(open(path).read().split("\n"), 1):
                m = re.match(r"^(template.*)?\s*__(device|global)__.*?\b([A-Za-z_0-9]+)\s*\(", l)
                if m and not l.startswith(" "): starts.append((i, m.group(3)))
        except OSError:
            pass
        srcs[path] = starts
    name = "?"
    for s, nm in srcs[path]:
        if s <= n: name = nm
        else: break
    return name
inside = False; cur = ("?", 0); agg = collections.Counter(); total = 0
for l in open(sass):
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        f = os.path.basename(cur[0]); fn = func_of(cur[0], cur[1])
        agg[(f, fn)] += 1; total += 1
print("kernel %s: %d instructions = %.1f KB" % (kern, total, total * 16 / 1024))
for (f, fn), n in agg.most_common(40):
    print("%6d  %5.1f%%  %6.1f KB  %s:%s" % (n, 100.0 * n / total, n * 16 / 1024, f, fn))
