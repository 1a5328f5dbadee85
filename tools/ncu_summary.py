#!/usr/bin/env python
"""Summarises the ncu captures of a build into profiles/<prefix>_ncu_summary.json (read by bench.py) and a text table:

  python tools/ncu_summary.py --launches gpurun_out/r02_launches.csv.gz --rep gpurun_out/r02_steady.ncu-rep \
         --rep gpurun_out/r02_side.ncu-rep --events 2000000 --prefix profiles/r02

--launches: CSV of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` over one
pipeline pass of --events events (serialised, cold-cache: the kernels' SHARES of the pass are meaningful, not the absolutes).
--rep: `ncu --set full` reports; per captured launch the keys the roofline discussion needs.
"""
import argparse
import collections
import csv
import gzip
import io
import json
import subprocess

KEYS = [("gpu__time_duration.sum", "time_us"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
        ("dram__bytes_read.sum", "dram_read_bytes"), ("dram__bytes_write.sum", "dram_write_bytes"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__inst_executed.sum", "inst"), ("smsp__inst_executed_pipe_fp64.sum", "fp64_inst"),
        ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_load_sectors"),
        ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "l1_sectors_per_global_load_request"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def read_launches(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(io.TextIOWrapper(op(path, "rb"), errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: {"launches": 0, "time_ms": 0.0, "dram_bytes": 0.0})
    ids = set()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        k = r[ki].split("(")[0].replace("void ", "").replace("art::", "")
        k = k.split("<")[0]
        v = num(r[vi])
        if v is None:
            continue
        unit = r[ui].lower()
        if r[mi] == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
            agg[k]["time_ms"] += v * scale
            if (r[0], k) not in ids:
                ids.add((r[0], k))
                agg[k]["launches"] += 1
        elif r[mi].startswith("dram__bytes"):
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            agg[k]["dram_bytes"] += v * scale
    return dict(agg)


def read_rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"kernel": r[ki].split("(")[0].replace("void ", "").replace("art::", "")}
        for name, short in KEYS:
            if name in hdr:
                i = hdr.index(name)
                v = num(r[i])
                if v is not None and short == "time_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[i].lower(), 1.0)
                if v is not None and short.endswith("_bytes"):
                    v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[i].lower(), 1.0)
                d[short] = v
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--events", type=int, required=True)
    ap.add_argument("--prefix", default="profiles/r02")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    summary = {"source": {"launches": a.launches, "reports": a.rep, "events_of_the_profiled_pass": a.events, "note": a.note}}
    lines = []
    if a.launches:
        L = read_launches(a.launches)
        tot = sum(v["time_ms"] for v in L.values())
        dram = sum(v["dram_bytes"] for v in L.values())
        summary["launch_list"] = {k: {**v, "share_of_kernel_time": v["time_ms"] / tot} for k, v in sorted(L.items(), key=lambda kv: -kv[1]["time_ms"])}
        summary["kernel_time_ms_of_the_pass_serialised"] = tot
        summary["dram_bytes_of_the_pass"] = dram
        summary["dram_bytes_per_step"] = dram * 1_000_000 / a.events  # per 1e6-event step
        lines.append(f"{sum(v['launches'] for v in L.values())} launches, {tot:.1f} ms of kernel time (serialised under ncu), "
                     f"{dram / 1e9:.2f} GB of DRAM traffic for {a.events} events")
        for k, v in summary["launch_list"].items():
            lines.append(f"  {k:22s} {v['launches']:6d} launches {v['time_ms']:9.2f} ms {100 * v['share_of_kernel_time']:5.1f} %  "
                         f"{v['dram_bytes'] / 1e9:8.2f} GB  {v['dram_bytes'] / max(v['time_ms'], 1e-9) / 1e6:8.1f} GB/s")
    caps = []
    for rp in a.rep:
        caps += read_rep(rp)
    summary["captures"] = caps
    for c in caps:
        lines.append("---- " + c["kernel"])
        lines.append("  " + "  ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in c.items() if k != "kernel"))
    json.dump(summary, open(a.prefix + "_ncu_summary.json", "w"), indent=1)
    open(a.prefix + "_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
