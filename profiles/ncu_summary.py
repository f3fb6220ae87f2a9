"""Summarise an ncu --page raw --csv dump: python profiles/ncu_summary.py <raw.csv> [pattern...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "launch__registers_per_thread", "launch__occupancy_limit", "sm__throughput.avg.pct",
                        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
                        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
                        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__average_warp", "smsp__warps_issue_stalled",
                        "smsp__inst_executed.sum", "launch__grid_size", "sm__cycles_elapsed.max", "local"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            print("  %-95s %s %s" % (h, r[i], units[i]))
