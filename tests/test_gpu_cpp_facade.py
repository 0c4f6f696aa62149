"""The boundary from a compiled host: build tests/cpp/facade_smoke.cpp against include/matrel.hpp + the in-tree .so
with g++ (no nvcc, no torch, no Python in the process) and run it on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_facade_end_to_end(tmp_path):
    exe = str(tmp_path / "facade_smoke")
    libdir = os.path.join(ROOT, "matrel_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_smoke.cpp"),
                        "-L", libdir, "-lmatrel_b200", f"-Wl,-rpath,{libdir}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK facade_smoke" in r.stdout, r.stdout + r.stderr
