"""Per-source-line stall samples / executed instructions of one kernel from an .ncu-rep (needs -lineinfo builds).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [FUNCTION_SUBSTRING] [--top N]

Runs `ncu -i REPORT --page source --print-source cuda,sass --csv` and sums, per (file, line), the warp stall samples,
warp instructions and thread instructions of the SASS attributed to it.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    fsub = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg, fpath, fname, hdr = {}, None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Function Name":
            fname = r[1]
        elif r[0] == "Line No":
            hdr = r
            iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        elif hdr and r[0] != "" and (fsub is None or fsub in fname):
            try:
                key = (fpath.split("/")[-1], int(r[0]), r[1].strip())
                s, i, t = int(r[iS]), int(r[iI]), int(r[iT])
            except ValueError:
                continue
            a = agg.setdefault(key, [0, 0, 0])
            a[0] += s; a[1] += i; a[2] += t
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print(f"total samples {ts}  warp instructions {ti}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0][:22]:22s}:{key[1]:4d} {a[0] * 100 / ts:5.1f}% smp {a[1] * 100 / ti:5.1f}% inst "
              f"thr/inst {a[2] / max(a[1], 1):4.1f} | {key[2][:100]}")


if __name__ == "__main__":
    main()
