#!/usr/bin/env python
"""Copy one GPU round (tools/gpu_round.sh <tag>) from gpurun_out/ into profiles/: bench lines, ncu launch list, summaries of
the ncu --set full capture (raw metrics + hottest lines, samples by function), clocks summary, profiles/ncu_traffic.json.
usage: python tools/collect_round.py <tag>"""
import csv, glob, io, json, os, statistics, subprocess, sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for f in sorted(glob.glob(os.path.join(G, tag + "_bench*.json"))):
    lines = [l for l in open(f) if l.startswith("{")]
    if lines:
        open(os.path.join(P, os.path.basename(f)), "w").write(lines[-1])
src = os.path.join(G, tag + "_launches_ctrl4096.csv")
if os.path.exists(src):
    open(os.path.join(P, os.path.basename(src)), "w").write(open(src).read())
rep = os.path.join(G, tag + "_prof_ctrl4096.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "60"], capture_output=True, text=True).stdout
    open(os.path.join(P, tag + "_ncu_h8t_ctrl4096.txt"), "w").write(out)
    byf = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), rep, os.path.join(ROOT, "autonomous-racing-lpv-mpp-mpc_b200", "csrc", "lpv_h8t.cuh")], capture_output=True, text=True).stdout
    open(os.path.join(P, tag + "_ncu_h8t_by_function.txt"), "w").write(byf)
    rd = wr = None
    for l in out.splitlines():
        if l.startswith("dram__bytes_read.sum "): rd = l.split()
        if l.startswith("dram__bytes_write.sum "): wr = l.split()
    if rd and wr:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = float(rd[2].replace(",", "")) * scale[rd[1]] + float(wr[2].replace(",", "")) * scale[wr[1]]
        tf = os.path.join(P, "ncu_traffic.json")
        t = json.load(open(tf)) if os.path.exists(tf) else {}
        t["ctrl4096"] = {"variant": 6, "dram_bytes_per_launch": int(tot),
                         "source": "profiles/%s_ncu_h8t_ctrl4096.txt (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, one launch)" % tag}
        json.dump(t, open(tf, "w"), indent=1)
        print("dram bytes per launch", int(tot))
clk = os.path.join(G, tag + "_clocks.csv")
if os.path.exists(clk):
    sm, reasons = [], set()
    for r in csv.reader(open(clk)):
        if len(r) >= 5 and r[1].strip().split()[0].isdigit():
            sm.append(int(r[1].strip().split()[0])); reasons.add(r[4].strip())
    if sm:
        open(os.path.join(P, tag + "_clocks_summary.txt"), "w").write(
            "nvidia-smi during tools/gpu_round.sh %s (500 ms samples): n=%d, clocks.sm min/median/max = %d/%d/%d MHz, clocks_event_reasons.active values seen: %s\n"
            % (tag, len(sm), min(sm), statistics.median(sm), max(sm), sorted(reasons)))
print("\n".join(sorted(os.path.basename(f) for f in glob.glob(os.path.join(P, tag + "_*")))))
