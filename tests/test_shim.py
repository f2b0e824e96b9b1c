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
