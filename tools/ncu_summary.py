#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md argues with.  usage: ncu_summary.py prof.ncu-rep [out.json]"""
import csv, io, json, re, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__sectors_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_requests_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_config_size", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "sm__maximum_warps_per_active_cycle_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for r in rows[2:]:
    d = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            d[w] = (r[i] + " " + units[i]).strip()
    for i, h in enumerate(hdr):
        if re.search(r"issue_stalled.*per_issue_active\.ratio", h) and "not_issued" not in h:
            try:
                if float(r[i]) > 0.5:
                    d[h] = r[i]
            except ValueError:
                pass
    out.append(d)
txt = json.dumps(out[0] if len(out) == 1 else out, indent=1)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
