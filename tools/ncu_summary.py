"""Key metrics of one kernel launch out of an `ncu --set full` report, as JSON (what profiles/*_ncu_full_summary.json hold).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu_full_summary.json"""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {"report": rep.split("/")[-1]}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            res["kernel"] = v
        if h in KEEP or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                res[h] = {"value": float(v), "unit": u}
            except ValueError:
                res[h] = {"value": v, "unit": u}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
