"""CPU check of the register-only rectangle intersection (csrc/mpp_clip.cuh, the routine behind RectangleOverlapEnergy,
prior_energies.py:12-24): the header is plain C++ as well, so tools/clip_check.cu is compiled with g++ and compares it with
an independent float64 Sutherland-Hodgman clip on random, class-aligned, degenerate and near-degenerate rectangle pairs."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_quad_box_area_against_float64_clip(tmp_path):
    exe = str(tmp_path / "clip_check")
    subprocess.run(["g++", "-O2", "-x", "c++", "-o", exe, os.path.join(ROOT, "tools", "clip_check.cu")], check=True)
    out = subprocess.run([exe, "10"], capture_output=True, text=True)
    sys.stdout.write(out.stdout[-600:])
    assert out.returncode == 0, out.stdout[-2000:]
    assert "worst |area error|" in out.stdout
