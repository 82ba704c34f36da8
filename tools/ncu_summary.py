"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ cite.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = rows[0]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"== {path} :: {d.get('Kernel Name', '?')[:70]}")
            for k in KEYS:
                if k in d and d[k] != "":
                    print(f"  {k} = {d[k]}")
            st = {k[len(STALL):]: float(v) for k, v in d.items() if k.startswith(STALL) and not k.endswith("_not_issued") and v}
            tot = sum(st.values()) or 1.0
            top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
            print("  stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in top))


if __name__ == "__main__":
    main()
