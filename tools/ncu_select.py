"""python tools/ncu_select.py <file.ncu-rep> -> the metrics the roofline / stall discussion uses, one per line"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__cluster", "launch__occupancy_limit", "launch__registers_per_thread", "sass__inst_executed_local", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct", "sm__throughput.avg.pct", "sm__warps_active.avg.pct",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic", "dram__throughput.avg.pct")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keep) and not h.endswith(".peak_sustained") and "per_second" not in h:
        print("%-90s %-14s %s" % (h, u, v))
