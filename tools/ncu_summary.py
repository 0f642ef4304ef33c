"""Summarise an ncu report (raw page CSV) for the step kernel: python tools/ncu_summary.py <rep> [kernel-substr]"""
import csv, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
data = [r for r in data if pat in r[4]]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy", "sm__pipe_fmalite", "sm__pipe_xu_cycles_active", "sm__pipe_tensor_cycles_active", "sm__ops_path_tensor_src_bf16_dst_fp32", "sm__inst_executed_pipe_tmem", "sass__inst_executed_local",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "warp_issue_stalled", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__sass_inst_executed_op_local", "local_load", "local_store", "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]
for i, h in enumerate(hdr):
    if any(k in h for k in KEYS):
        vals = [r[i] for r in data]
        print(f"{h} [{units[i]}] = {vals}")
