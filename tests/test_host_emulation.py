"""The arithmetic of CUDA sources checked WITHOUT a GPU: device functions of sac_b200/csrc/sparse.cu and model_dev.cuh
(map coder model step, contexts, range coder, O(1) rank mapping) are compiled for the host (tests/host_emul/) and compared
with the CPU restatement on the CPU. Needs nvcc (present in the build image and on the GPU box)."""
import os
import shutil
import subprocess

import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sparse_device_code_on_the_host_equals_the_oracle(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    ol.oracle()                                   # builds oracle/liboracle.so if needed
    exe = str(tmp_path / "sparse_emul")
    cmd = [nvcc, "-std=c++17", "-O1", "-arch=sm_100a", "--fmad=false", "-w", "-I", os.path.join(ROOT, "sac_b200", "csrc"), "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_emul", "sparse_emul.cu"), "-o", exe, "-L",
           os.path.join(ROOT, "oracle"), "-loracle", "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")]
    subprocess.run(cmd, check=True, capture_output=True, timeout=600)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL EQUAL" in r.stdout, r.stdout + r.stderr
