"""SASS instructions attributed to one source line of one kernel, with their stall samples by reason.

    python tools/ncu_sass_of_line.py REPORT.ncu-rep KERNEL_REGEX FUNCTION_SUBSTRING FILE_SUBSTRING LINE [LINE2]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kre, fsub, file_sub = sys.argv[1:5]
    lo = int(sys.argv[5])
    hi = int(sys.argv[6]) if len(sys.argv) > 6 else lo
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
    fpath = fname = hdr = None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Function Name":
            fname = r[1]
        elif r[0] == "Line No":
            hdr = r
            stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        elif hdr and fsub in fname and file_sub in fpath:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            if lo <= ln <= hi and int(r[iS] or 0) >= 0:
                reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:3]
                print(f"{ln:5d} smp {int(r[iS]):5d} inst {int(r[iI]):9d} thr {int(r[iT]) / max(int(r[iI]), 1):4.1f} | {r[3][:70]:70s} | "
                      + " ".join(f"{n}:{c}" for c, n in reasons if c))


if __name__ == "__main__":
    main()
