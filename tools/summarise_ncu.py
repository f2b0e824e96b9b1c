"""Condenses an .ncu-rep (ncu --set full) into a small CSV of the metrics the design decisions rest on.

    python tools/summarise_ncu.py gpurun_out/prof.ncu-rep profiles/r02_xxx.csv
    python tools/summarise_ncu.py --profile gpurun_out/prof.ncu-rep profiles/r02_profile.json
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


def profile_json(rep, out):
    """Per kernel (bare name; the largest launch of each kernel in the report): DRAM bytes (read + write), warp
    instructions executed, duration, issue-slot utilisation -- what bench.py quotes as roofline.traffic and
    roofline.executed (profiles/r02_profile.json)."""
    import json
    import re
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "Kernel Name", "smsp__inst_executed.sum",
                                     "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active")}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    res = {}
    for r in rows[2:]:
        m = re.search(r"(\w+_kernel)", r[col["Kernel Name"]])
        name = m.group(1) if m else r[col["Kernel Name"]]
        ir, iw = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
        total = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        inst = float(r[col["smsp__inst_executed.sum"]])
        it = col["gpu__time_duration.sum"]
        rec = {"dram_bytes": int(total), "inst_executed": int(inst), "ms_under_ncu": float(r[it]) * tscale.get(units[it], 1.0),
               "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
               "kernel": r[col["Kernel Name"]][:80]}
        if name not in res or inst > res[name]["inst_executed"]:   # (e.g. the family instance of an integer-keypoint batch is empty)
            res[name] = rec
    with open(out, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "--profile":
        profile_json(sys.argv[2], sys.argv[3])
    else:
        main(sys.argv[1], sys.argv[2])
