"""Key metrics of the first kernel in an .ncu-rep as one JSON object (the numbers profiles/README.md quotes).
usage: python tools/ncu_summary.py file.ncu-rep"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration_us", "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct", "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_inst",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks", "smsp__inst_executed_op_shared_ld.sum": "lds_instructions", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
}
STALL = "smsp__average_warps_issue_stalled_"


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out, stalls = {"kernel": None}, {}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            out["kernel"] = v
        if h in WANT:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in ("Mbyte", "MB"):
                x *= 1e6
            if u in ("Kbyte", "KB"):
                x *= 1e3
            if u in ("Gbyte", "GB"):
                x *= 1e9
            if u == "ms":
                x *= 1e3
            if u == "ns":
                x *= 1e-3
            out[WANT[h]] = x
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls[h[len(STALL):-len("_per_issue_active.ratio")]] = float(v)
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1.0
    out["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.01}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
