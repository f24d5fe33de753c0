"""key counters of one kernel launch from an ncu report (ncu -i <rep> --page raw --csv), as committed under profiles/
usage: python scripts/ncu_summary.py <report.ncu-rep> [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sass__inst_executed_register_spilling", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if want and want not in name:
        continue
    print("kernel:", name[:140])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h), reverse=True)
    print("  top stalls per issue: " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in stalls[:7]))
    break
