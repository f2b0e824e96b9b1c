"""The C++ drop-in header (include/legoslam_gpu/algorithm_shim.h): compiles without OpenCV (CPU test) and,
on the GPU box, reproduces the oracle through the reference's own entry-point signature."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "shim", "shim_test.cpp")
EXE = os.path.join(ROOT, "tests", "shim", "shim_test")


def _build():
    from lego_slam_b200 import build
    from oracle import binding
    lib = build.build()
    ora = binding.build()
    cmd = ["g++", "-std=c++11", "-O2", "-o", EXE, SRC, lib, ora,
           f"-Wl,-rpath,{os.path.dirname(lib)}", f"-Wl,-rpath,{os.path.dirname(ora)}"]
    subprocess.check_call(cmd)


def test_shim_header_compiles_and_links_without_opencv():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_shim_matches_oracle_on_gpu():
    _build()
    res = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr


# ---- the integration inside the reference's own tree: its declarations (default arguments included) visible ----
SRC_CV = os.path.join(ROOT, "tests", "shim", "shim_opencv_test.cpp")
EXE_CV = os.path.join(ROOT, "tests", "shim", "shim_opencv_test")


def _build_cv():
    from lego_slam_b200 import build
    lib = build.build()
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "shim", "fake_opencv"),
           "-o", EXE_CV, SRC_CV, lib, f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.check_call(cmd)


def test_entry_point_definitions_compile_against_the_reference_prototypes():
    """LEGOSLAM_GPU_DEFINE_ENTRY_POINTS with the reference's prototypes (defaults included) and an OpenCV look-alike in
    one translation unit: well-formed C++ (no default argument given twice), links, and without a GPU reports the
    missing device instead of falling back."""
    _build_cv()
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([EXE_CV], capture_output=True, text=True, timeout=60)
        assert res.returncode == 3 and "lego_klt_create" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_entry_point_definitions_run_on_gpu():
    _build_cv()
    res = subprocess.run([EXE_CV], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
