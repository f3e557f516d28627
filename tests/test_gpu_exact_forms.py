"""Device side of tests/test_exact_forms_cpu.py: the check-free quotient of the window sampler (r_div_nocheck, mpp_device.cuh) and
the FMA-corrected /3 of shape_terms (mpp_sweep2.cuh) equal the IEEE quotients bit for bit on 16.7 M operand pairs.  The checker
(tools/div_check.cu) includes the product header, so it exercises the very function the kernels inline."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_check_free_quotients_equal_ieee_division(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "divcheck")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, os.path.join(ROOT, "tools", "div_check.cu")], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "IEEE division: 0 mismatches" in out.stdout and "__fdiv_rn: 0 mismatches" in out.stdout, out.stdout
