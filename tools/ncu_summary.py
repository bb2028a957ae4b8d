#!/usr/bin/env python
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics we track."""
import csv, subprocess, sys, io
KEEP = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed_op_shared_atom.sum','smsp__inst_executed_op_global_red.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__sass_average_branch_targets_threads_uniform.pct','sm__sass_thread_inst_executed_op_dfma_pred_on.sum','sm__sass_thread_inst_executed_op_dadd_pred_on.sum','sm__sass_thread_inst_executed_op_dmul_pred_on.sum','sm__sass_thread_inst_executed_op_ffma_pred_on.sum','sm__sass_thread_inst_executed_op_fadd_pred_on.sum','sm__sass_thread_inst_executed_op_fmul_pred_on.sum','smsp__sass_thread_inst_executed_op_fp64_pred_on.sum','smsp__sass_thread_inst_executed_op_fp32_pred_on.sum','smsp__sass_thread_inst_executed_op_integer_pred_on.sum','smsp__thread_inst_executed.sum','smsp__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
def main(path, out=None, title=""):
    raw = subprocess.run(["ncu","-i",path,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title or path}", "kernel,metric,unit,value"]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        for h,u,v in zip(hdr,units,vals):
            if h in KEEP or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
                lines.append(f"{name.split('(')[0].replace(',', ';')},{h},{u},{v}")
    txt = "\n".join(lines)+"\n"
    if out: open(out,"w").write(txt)
    print(txt)
if __name__ == "__main__":
    main(*sys.argv[1:])
