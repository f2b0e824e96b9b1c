"""Condenses an .ncu-rep (ncu --set full) into a small CSV of the metrics the design decisions rest on.

    python tools/summarise_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx.csv
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"))]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [hdr[i].replace(STALLS, "stall_").replace("_per_issue_active.ratio", "") + (" [" + units[i] + "]" if units[i] else "") for i in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index("Kernel Name")][:60]] + [r[i] for i in cols])


def traffic_json(rep, out):
    """DRAM bytes per launch (read + write) of every kernel in the report, keyed by the bare kernel name: the
    numbers bench.py quotes as roofline.traffic (first launch of each kernel)."""
    import json
    import re
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        m = re.search(r"(\w+_kernel)", r[ik])
        name = m.group(1) if m else r[ik]
        total = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        if name not in res or total > res[name]:      # (klt_lane_kernel<.., true> is the empty FAMILIES instance)
            res[name] = total
    with open(out, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic_json(sys.argv[2], sys.argv[3])
    else:
        main(sys.argv[1], sys.argv[2])
