"""Summarise an .ncu-rep: headline raw metrics and the top stall lines of the source page.  usage: python scripts/ncu_quick.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__pipe_tensor_cycles_active_realtime.avg.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum ", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum ", "launch__registers_per_thread ", "sm__throughput.avg.pct",
        "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg ", "sm__warps_active.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "smsp__inst_executed.sum ", "sm__inst_executed_pipe_fp64", "smsp__issue_active.avg.pct", "sm__pipe_fp64_cycles_active"]
for h, u, v in zip(hdr, units, vals):
    if any((h + " ").startswith(w) or w.strip() == h for w in want):
        print("%-90s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data if len(r) > isamp and r[isamp].isdigit())
print("total samples", tot)
top = sorted([(int(r[isamp]), i) for i, r in enumerate(data) if len(r) > isamp and r[isamp].isdigit()], reverse=True)[:ntop]
for s, i in sorted(top, key=lambda x: x[1]):
    r = data[i]
    st = sorted(((h, int(r[hdr.index(h)])) for h in stalls if r[hdr.index(h)].isdigit() and int(r[hdr.index(h)]) > 0), key=lambda x: -x[1])[:2]
    print("%5d %7d %9s  %-80s %s" % (i, s, r[iex], r[ia].strip()[:80], st))
