#!/usr/bin/env python
"""SASS evidence of the built library: per kernel, how often the instructions that carry the design appear (cuobjdump -sass of
visualslam_android_b200/libvslam_b200.so), plus a few lines of context for the first occurrence of each.
    python profiles/tools/sass_evidence.py > profiles/<tag>_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "visualslam_android_b200", "libvslam_b200.so")
WHAT = [("UBLKCP", "cp.async.bulk global->shared (TMA engine): strip staging of the FAST kernels"),
        ("SYNCS", "mbarrier arrive / try_wait (completion of the bulk copy)"),
        ("VABSDIFF4", "byte-SIMD |a-b| on four pixels: FAST rejection test"),
        ("IDP.4A", "dp4a: half-sampling (pyramid), ZMSSD sums"),
        ("LDS.128", "128-bit shared-memory loads: 16 pixels per lane"),
        ("RED.E.OR|REDG.E.OR|RED.E.OR.STRONG", "fire-and-forget atomic OR into the corner bitmask"),
        ("REDUX", "warp reduce in one instruction"),
        ("SHFL", "warp shuffles (scans, butterflies)"),
        ("VOTE|VOTEU", "ballots (queue compaction)"),
        ("DADD|DMUL|DFMA", "FP64 (projection, WLS)"),
        ("UTC|HMMA|HGMMA", "tensor-core instructions (none expected: nothing here is a dense contraction)")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern = None; body = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        ks = re.findall(r"\d+(k_[a-z][a-z0-9_]*?)(?=ILi|E|$)", m.group(1)); kern = (ks[-1] if ks else m.group(1)[:40]) + ("" if "ILi" not in m.group(1) else "<" + re.search(r"ILi(\d+)", m.group(1)).group(1) + ">")
        body.setdefault(kern, []); continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        body[kern].append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; instruction counts are static occurrences per kernel)\n")
print("kernel".ljust(26) + "".join(p.split("|")[0].ljust(11) for p, _ in WHAT) + "total")
for k, ins in body.items():
    if not k.startswith("k_"): continue
    row = k.ljust(26)
    for pat, _ in WHAT:
        n = sum(1 for i in ins if re.search(r"\b(" + pat.replace(".", r"\.") + r")", i))
        row += str(n).ljust(11)
    print(row + str(len(ins)))
print()
for pat, why in WHAT:
    print(f"## {pat}: {why}")
    shown = 0
    for k, ins in body.items():
        if not k.startswith("k_"): continue
        for idx, i in enumerate(ins):
            if re.search(r"\b(" + pat.replace(".", r"\.") + r")", i):
                print(f"  {k}:"); [print("    " + x.strip()) for x in ins[max(0, idx - 1):idx + 2]]; shown += 1; break
        if shown >= 2: break
    if not shown: print("  (none)")
