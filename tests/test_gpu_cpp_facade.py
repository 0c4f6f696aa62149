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


def test_c_abi_grid_without_python(tmp_path):
    """one process drives the GPUs of the box through the C ABI alone (mr_init_grid / mr_dmatrix_*): multiply with
    peer pulls, element-wise, NCCL reduction, re-partitioning.  Uses 2 GPUs when the box has them, else the 1 x 1 grid."""
    exe = str(tmp_path / "grid_smoke")
    libdir = os.path.join(ROOT, "matrel_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "grid_smoke.cpp"),
                        "-L", libdir, "-lmatrel_b200", f"-Wl,-rpath,{libdir}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch
    for ngpus in sorted({1, min(2, torch.cuda.device_count()), min(4, torch.cuda.device_count())}):
        r = subprocess.run([exe, str(ngpus)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and f"OK grid_smoke gpus={ngpus}" in r.stdout, r.stdout + r.stderr
