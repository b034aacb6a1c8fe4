"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag> [--rep gpurun_out/<tag>_bp.ncu-rep] [--launches gpurun_out/<tag>_launches.csv]
"""
import argparse, collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    agg = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        key = (d["Kernel Name"].split("(")[0][:90], d["Block Size"], d["Grid Size"])
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("kernel,block,grid,launches,total_us,max_us,share_pct\n")
        for (k, b, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",\"{b}\",\"{g}\",{a[0]},{a[1]:.1f},{a[2]:.1f},{100 * a[1] / tot:.2f}\n")
    print(open(out).read())


def rep(path, out):
    # path: the .ncu-rep, or the raw-page CSV scripts/profile.sh exports on the box when the report is too big to travel
    raw = open(path).read() if path.endswith(".csv") else \
        subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d["Kernel Name"], "block": d["Block Size"], "grid": d["Grid Size"]}
        for k in KEEP:
            if k in d and d[k] != "":
                e[k] = d[k] + " " + units[hdr.index(k)]
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lp = a.launches or os.path.join(ROOT, "gpurun_out", a.tag + "_launches.csv")
    rp = a.rep or os.path.join(ROOT, "gpurun_out", a.tag + "_bp.ncu-rep")
    if os.path.exists(lp):
        launches(lp, os.path.join(ROOT, "profiles", a.tag + "_launches_summary.csv"))
    if not os.path.exists(rp) and os.path.exists(os.path.join(ROOT, "gpurun_out", a.tag + "_bp_raw.csv")):
        rp = os.path.join(ROOT, "gpurun_out", a.tag + "_bp_raw.csv")
    if os.path.exists(rp):
        rep(rp, os.path.join(ROOT, "profiles", a.tag + "_ncu_full_summary.json"))
