"""Compact summary of an ncu report (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep [...]"""
import csv, subprocess, sys, io, json

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]

def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, row):
            if h == "Kernel Name":
                d["kernel"] = v.split("(")[0]
            if h in KEYS:
                d[h] = f"{v} {u}".strip()
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    fv = float(v)
                except ValueError:
                    continue
                if fv >= 3.0:
                    d.setdefault("stalls_pct", {})[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")] = fv
        res.append(d)
    return res

if __name__ == "__main__":
    for p in sys.argv[1:]:
        for d in summarize(p):
            print(json.dumps(d, indent=1))
