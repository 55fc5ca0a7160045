#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report into the summaries kept under profiles/.

    python profiles/tools/ncu_extract.py REPORT.ncu-rep --tag r02a [--streams 256 --frame 640 480] [--kernels k_pyramid_fast k_fast_levels ...]

Writes  profiles/<tag>_ncu_full_summary.json   per kernel launch: duration, DRAM bytes, warp-instructions, issue utilisation, occupancy, registers, ...
        profiles/<tag>_ncu_counters.json       what bench.py reads for roofline.traffic / issue_roofline (first launch of every kernel name)
        profiles/<tag>_hot_<kernel>.txt        per source line and per opcode: executed warp-instructions, stall samples, active threads
Needs only the `ncu` CLI (no GPU)."""
import argparse, collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
RAW = {
    "duration_us": ("gpu__time_duration.sum", 1.0),
    "dram_bytes_read": ("dram__bytes_read.sum", None),
    "dram_bytes_write": ("dram__bytes_write.sum", None),
    "warp_instructions": ("smsp__inst_executed.sum", 1.0),
    "threads_per_instruction": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
    "issue_slot_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "registers_per_thread": ("launch__registers_per_thread", 1.0),
    "grid": ("launch__grid_size", 1.0),
    "block": ("launch__block_size", 1.0),
    "dyn_smem_bytes": ("launch__shared_mem_per_block_dynamic", None),
    "static_smem_bytes": ("launch__shared_mem_per_block_static", None),
    "lsu_data_pipe_pct": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l2_throughput_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l1_hit_pct": ("l1tex__t_sector_hit_rate.pct", 1.0),
    "local_load_sectors": ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", 1.0),
    "local_store_sectors": ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", 1.0),
    "stall_barrier_per_issue": ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0),
    "stall_long_scoreboard_per_issue": ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    "stall_short_scoreboard_per_issue": ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1.0),
    "stall_mio_throttle_per_issue": ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1.0),
    "stall_wait_per_issue": ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0),
    "stall_not_selected_per_issue": ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", 1.0),
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def short(name):
    m = re.search(r"(k_[A-Za-z0-9_]+)", name)
    return m.group(1) if m else name[:40]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def raw_summary(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        k = {"kernel": short(r[hdr.index("Kernel Name")]), "id": int(r[hdr.index("ID")])}
        for key, (metric, _) in RAW.items():
            if metric not in hdr:
                continue
            i = hdr.index(metric)
            v = num(r[i])
            if v is None:
                continue
            u = units[i]
            if u in UNIT_SCALE:
                v *= UNIT_SCALE[u]
            k[key] = v
        out.append(k)
    return out


def hot_lines(rep, kernel, top=28):
    txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel + "$|" + kernel + r"\("])
    # several launches of the same kernel follow one another, each with its own header: keep the first (largest grid comes first in our captures)
    lines = collections.OrderedDict(); ops = collections.Counter(); ops_samp = collections.Counter()
    hdr = None; cur_file = ""; seen_fn = 0
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            if hdr is not None and cur_file == first_file and lines:
                seen_fn += 1
            hdr = r
            if seen_fn == 0 and not lines:
                first_file = cur_file
            i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples"); i_thr = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0] != "":                       # a source line: aggregated over its SASS
            key = (cur_file, int(r[0]))
            inst, samp, thr = num(r[i_inst]) or 0.0, num(r[i_samp]) or 0.0, num(r[i_thr]) or 0.0
            if key in lines:
                a = lines[key]; a[1] += inst; a[2] += samp; a[3] += thr
            else:
                lines[key] = [r[1].strip(), inst, samp, thr]
        else:                                # a SASS line
            op = r[3].strip().split()
            if op and op[0].startswith("@"):
                op = op[1:]
            if op:
                name = op[0].split(".")[0]
                ops[name] += num(r[i_inst]) or 0.0; ops_samp[name] += num(r[i_samp]) or 0.0
    tot_i = sum(v[1] for v in lines.values()) or 1.0
    tot_s = sum(v[2] for v in lines.values()) or 1.0
    out = [f"{kernel}: total warp-inst {tot_i:.0f}, stall samples {tot_s:.0f} (all launches of the capture)", "--- by instructions"]
    fmt = lambda k, v: f"  {100 * v[1] / tot_i:4.1f}% inst  {100 * v[2] / tot_s:4.1f}% samp thr/inst {v[3] / max(v[1], 1.0):4.1f}  {k[0]}:{k[1]}  {v[0][:110]}"
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(fmt(k, v))
    out.append("--- by samples")
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
        out.append(fmt(k, v))
    out.append("--- opcodes")
    to = sum(ops.values()) or 1.0; ts = sum(ops_samp.values()) or 1.0
    for name, c in ops.most_common(18):
        out.append(f"{name:12s} {100 * c / to:5.1f}% inst  {100 * ops_samp[name] / ts:5.1f}% samp")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report"); ap.add_argument("--tag", required=True)
    ap.add_argument("--streams", type=int, default=256); ap.add_argument("--frame", type=int, nargs=2, default=[640, 480])
    ap.add_argument("--kernels", nargs="*", default=None); ap.add_argument("--command", default="python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline")
    ap.add_argument("--counters-name", default=None, help="file name of the counters json (default <tag>_ncu_counters.json)")
    a = ap.parse_args()
    P = os.path.join(ROOT, "profiles")
    summ = raw_summary(a.report)
    json.dump({"command": a.command, "ncu": "--set full --clock-control none --import-source on -k regex:k_", "streams": a.streams, "frame": a.frame, "launches": summ},
              open(os.path.join(P, f"{a.tag}_ncu_full_summary.json"), "w"), indent=1)
    first = {}
    for k in summ:
        if k["kernel"] not in first or k.get("duration_us", 0) > first[k["kernel"]].get("duration_us", 0):
            first[k["kernel"]] = k          # the longest launch of a name (fine stage rather than coarse stage)
    json.dump({"source": f"profiles/{a.tag}_ncu_full_summary.json", "command": a.command, "streams": a.streams, "frame": a.frame, "kernels": first},
              open(os.path.join(P, a.counters_name or f"{a.tag}_ncu_counters.json"), "w"), indent=1)
    for kname in (a.kernels if a.kernels is not None else sorted(first)):
        open(os.path.join(P, f"{a.tag}_hot_{kname}.txt"), "w").write(hot_lines(a.report, kname))
    for k in summ:
        print(f"{k['kernel']:18s} {k.get('duration_us', 0):8.1f} us  inst {k.get('warp_instructions', 0) / 1e6:7.1f} M  issue {k.get('issue_slot_pct', 0):5.1f}%  warps {k.get('warps_active_pct', 0):5.1f}%  "
              f"regs {k.get('registers_per_thread', 0):.0f}  dram {((k.get('dram_bytes_read') or 0) + (k.get('dram_bytes_write') or 0)) / 1e6:7.1f} MB  thr/inst {k.get('threads_per_instruction', 0):.1f}")


if __name__ == "__main__":
    main()
