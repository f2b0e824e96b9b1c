"""SASS opcode histogram per kernel of a built library (cuobjdump -sass), for profiles/.

    python tools/sass_hist.py lego_slam_b200/liblego_klt.so [KERNEL_SUBSTRING ...] [--top N]

Prints, per kernel whose (demangled-ish) name contains one of the substrings (all kernels if none): the static
instruction count and the most frequent opcodes, with the Blackwell / Hopper-and-later markers the profiling guide
names (UTMALDG, UBLKCP, LDGSTS, FMUL2/FFMA2/FADD2, ...) listed explicitly even when they are not in the top N.
"""
import collections
import re
import subprocess
import sys

MARKERS = ["UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "FMUL2", "FFMA2", "FADD2", "DFMA", "F2F", "I2F", "F2I", "PRMT",
           "MOV", "IMAD", "LDS", "STS", "SYNCS", "ATOMS", "IDP", "SHF", "MUFU"]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    if "--top" in sys.argv:
        args = [a for a in args if a != sys.argv[sys.argv.index("--top") + 1]]
    lib, subs = args[0], args[1:]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, hist = None, None
    kernels = []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name, hist = m.group(1), collections.Counter()
            kernels.append((name, hist))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and hist is not None:
            hist[m.group(1)] += 1
    for name, hist in kernels:
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"legoklt::|\(anonymous namespace\)::|\(legoklt::PyramidView.*", "", short)
        if subs and not any(s in short for s in subs):
            continue
        total = sum(hist.values())
        marks = " ".join(f"{k}={hist[k]}" for k in MARKERS if hist.get(k))
        tops = " ".join(f"{k}={v}" for k, v in hist.most_common(top))
        print(f"{short}: {total} instructions\n    top: {tops}\n    markers: {marks}")


if __name__ == "__main__":
    main()
