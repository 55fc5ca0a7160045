#!/usr/bin/env python
"""Per source line: executed warp-instructions of one kernel of an ncu report, in source order (lines above a share threshold).
    python profiles/tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [pixels-or-units] [min_share]"""
import csv, io, subprocess, collections, sys
rep, kern = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
thresh = float(sys.argv[4]) if len(sys.argv) > 4 else 0.002
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
def num(s):
    try: return float(s.replace(",", ""))
    except ValueError: return 0.0
hdr = None; lines = collections.OrderedDict(); f = ""
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": f = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or r[0] == "" or r[0] == "Function Name": continue
    k = (f, int(r[0])); v = lines.setdefault(k, [r[1].strip(), 0.0, 0.0, 0.0]); v[1] += num(r[ii]); v[2] += num(r[ti]); v[3] += num(r[si])
tot = sum(v[1] for v in lines.values()); ts = sum(v[3] for v in lines.values()) or 1.0
print(f"{kern}: {tot:.0f} warp-instructions" + (f", {tot * 32 / units:.2f} lane-instructions per unit" if units else ""))
for k, v in sorted(lines.items()):
    if v[1] / tot > thresh or v[3] / ts > 0.01:
        per = f"{v[1] * 32 / units:6.2f}/unit " if units else ""
        print(f"{k[0][:16]:16s}:{k[1]:4d} {100 * v[1] / tot:5.2f}% inst {per}{100 * v[3] / ts:5.1f}% samp thr {v[2] / max(v[1], 1):4.1f}  {v[0][:110]}")
