"""Summarise an `ncu --set full` report (read here, no GPU needed) into a small JSON under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_c2.ncu-rep profiles/c2_kernel_ncu.json
    python tools/summarize_ncu.py gpurun_out/r02_ncu_stream_raw.csv profiles/r02_ncu_stream.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_requests_srcunit_tex.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):  # `ncu -i X.ncu-rep --page raw --csv` exported on the GPU box (gpurun brings back <= 64 MiB)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                try:
                    d[h] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    pass
        kernels.append(d)
    k = kernels[-1]

    def to_bytes(m):
        v, u = k[m]["value"], k[m]["unit"].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    summary = {"source": rep, "kernel": k["kernel"], "launches_in_report": len(kernels),
               "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
               "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
               "metrics": k}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps({x: summary[x] for x in ("kernel", "dram_bytes_per_launch", "dram_bytes_read", "dram_bytes_write")}))


if __name__ == "__main__":
    main()
