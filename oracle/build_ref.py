"""Builds oracle/_ref/: the reference's OWN hot-path translation unit, compiled here (TEST INFRASTRUCTURE ONLY).

    python oracle/build_ref.py            # needs /root/reference (the build container); outputs only into oracle/_ref/

`libklt_ref.so` = /root/reference/src/algorithm.cpp + /root/reference/include/legoslam/algorithm.h, both compiled
UNMODIFIED where they lie (nothing of the reference is copied into the repo), with the reference's own flags
(CMakeLists.txt:6-7: -std=c++11 -O3, no -march, no -ffast-math), against the stand-in headers in oracle/ref_stubs/
(OpenCV, Eigen, Sophus and glog are not installed in this image; the reference's own build system is not run).
What is NOT the reference's text in there: the cv:: types, cv::resize (= oracle/resize_u8.cpp, pinned to cv2),
the serial cv::parallel_for_, and Eigen's 2x2 pivoted LDLT (restated in ref_stubs/eigen_stub.hpp).

`libklt_ref_p{H}_l{L}.so` = the same text with THREE literals replaced through a sed-style substitution into a scratch
copy under oracle/_ref/ (git-ignored): half_patch_size (src/algorithm.cpp:40), pyramids (:135) and the scales table
(:137) -- for BASELINE configs C4 (5 levels) and C5 (11x11 patch), which generalise the reference's literals.
Those libraries report klt_ref_verbatim() == 0.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "ref_stubs")
REF = os.environ.get("LEGO_REFERENCE_ROOT", "/root/reference")
REF_TU = os.path.join(REF, "src", "algorithm.cpp")
REF_INC = os.path.join(REF, "include")
CXXFLAGS = ["-std=c++11", "-O3", "-fPIC", "-pthread", "-shared"]    # CMakeLists.txt:6-7 + what a .so needs
PARAM_VARIANTS = [(3, 5), (5, 4), (5, 3)]                              # (half_patch_size, pyramids)


def reference_present() -> bool:
    return os.path.isfile(REF_TU) and os.path.isfile(os.path.join(REF_INC, "legoslam", "algorithm.h"))


def lib_path(half_patch: int = 3, pyramids: int = 4) -> str:
    if (half_patch, pyramids) == (3, 4):
        return os.path.join(OUT, "libklt_ref.so")
    return os.path.join(OUT, f"libklt_ref_p{half_patch}_l{pyramids}.so")


def _compile(tu: str, out: str, defs: list[str]) -> None:
    cmd = ["g++"] + CXXFLAGS + defs + ["-I", STUBS, "-I", REF_INC, tu, os.path.join(STUBS, "ref_capi.cpp"),
                                      os.path.join(HERE, "resize_u8.cpp"), "-o", out + ".tmp"]
    subprocess.check_call(cmd)
    os.replace(out + ".tmp", out)


def _parametrised_tu(half_patch: int, pyramids: int) -> str:
    """Scratch copy of the reference TU with three literals replaced (see module docstring)."""
    text = open(REF_TU).read()
    subs = [
        (r"int half_patch_size = 3;", f"int half_patch_size = {half_patch};"),
        (r"int pyramids = 4;", f"int pyramids = {pyramids};"),
        (r"double scales\[\] = \{1\.0, 0\.5, 0\.25, 0\.125\};",
         "double scales[] = {1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 0.0078125};"),
    ]
    for pat, rep in subs:
        text, k = re.subn(pat, rep, text)
        if k != 1:
            raise RuntimeError(f"reference text changed: {pat!r} matched {k} times")
    # the scratch copy lives in a temporary directory OUTSIDE the repository and is deleted after the compile: no text of
    # the reference stays under this tree, not even in the git-ignored output directory
    import tempfile
    d = tempfile.mkdtemp(prefix="klt_ref_")
    path = os.path.join(d, f"algorithm_p{half_patch}_l{pyramids}.cpp")
    with open(path, "w") as f:
        f.write(text)
    return path


def build(force: bool = False) -> list[str]:
    """Builds every _ref library that is missing (all of them with force).  Returns the paths."""
    if not reference_present():
        raise RuntimeError(f"{REF} is not present: oracle/_ref can only be built in the build container")
    os.makedirs(OUT, exist_ok=True)
    built = []
    deps = [os.path.join(STUBS, f) for f in ("ref_capi.cpp", "eigen_stub.hpp", "opencv2/opencv.hpp",
                                             "legoslam/common_include.h")] + [os.path.join(HERE, "resize_u8.cpp"),
                                                                              os.path.abspath(__file__)]
    newest = max(os.path.getmtime(d) for d in deps)
    for hp, lv in [(3, 4)] + PARAM_VARIANTS:
        out = lib_path(hp, lv)
        if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
            if (hp, lv) == (3, 4):
                _compile(REF_TU, out, [])
            else:
                tu = _parametrised_tu(hp, lv)
                try:
                    _compile(tu, out, ["-DKLT_REF_PARAMETRISED", f"-DKLT_REF_HALF_PATCH={hp}", f"-DKLT_REF_PYRAMIDS={lv}"])
                finally:
                    import shutil
                    shutil.rmtree(os.path.dirname(tu), ignore_errors=True)
        built.append(out)
    return built


if __name__ == "__main__":
    for p in build(force="-f" in sys.argv):
        print(p)
