"""CPU: the intersection area behind RectangleOverlapEnergy (prior_energies.py:12-24) against two INDEPENDENT computations.

The reference delegates it to shapely 1.7.1 / GEOS 3.8.0 (absent here and on the GPU box), so the oracle, the stub that produced
the golden vectors and the CUDA clip are all restatements of convex clipping.  This file closes that circle as far as it can be
closed without shapely:

* against OpenCV's cv2.intersectConvexConvex (a different algorithm and code base, float32 points) on 10^4 random rectangle
  pairs: |difference| <= 3e-5 of the smaller area;
* against an EXACT rational clip (fractions.Fraction, no rounding at all) on a few hundred rectangles with integer corner
  coordinates (Pythagorean rotations), including identical, nested, edge-touching, corner-touching, collinear-edge and
  crossing pairs: the float64 oracle must agree to 1e-12 of the smaller area, the float32 device routine (csrc/mpp_clip.cuh,
  compiled for the host) to 1e-5.
The CUDA kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py)."""
import ctypes
import itertools
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from oracle import mpp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exact_area(poly):
    n = len(poly)
    s = Fraction(0)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        s += x0 * y1 - x1 * y0
    return abs(s) / 2


def _exact_clip_area(subject, clip):
    """Sutherland-Hodgman over Fractions (both rings convex): exact intersection area."""
    def signed(p):
        return sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p)))
    if signed(clip) < 0:
        clip = clip[::-1]
    out = list(subject)
    for i in range(len(clip)):
        if not out:
            return Fraction(0)
        ax, ay = clip[i]
        bx, by = clip[(i + 1) % len(clip)]
        ex, ey = bx - ax, by - ay
        inp, out = out, []
        for k in range(len(inp)):
            px, py = inp[k]
            qx, qy = inp[(k + 1) % len(inp)]
            sp = ex * (py - ay) - ey * (px - ax)
            sq = ex * (qy - ay) - ey * (qx - ax)
            if sp >= 0:
                out.append((px, py))
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                out.append((px + t * (qx - px), py + t * (qy - py)))
    return _exact_area(out) if len(out) >= 3 else Fraction(0)


def _int_rect(cx, cy, ux, uy, a, b):
    """Rectangle with integer corners: centre (cx, cy), half axes a * (ux, uy) and b * (-uy, ux)."""
    hx, hy, wx, wy = a * ux, a * uy, -b * uy, b * ux
    return [(Fraction(cx + hx + wx), Fraction(cy + hy + wy)), (Fraction(cx + hx - wx), Fraction(cy + hy - wy)),
            (Fraction(cx - hx - wx), Fraction(cy - hy - wy)), (Fraction(cx - hx + wx), Fraction(cy - hy + wy))]


def _exact_cases():
    dirs = [(1, 0), (0, 1), (1, 1), (3, 4), (4, 3), (5, 12), (8, 15), (-3, 4), (1, -1)]
    cases = []
    for (d1, d2) in itertools.product(dirs[:6], dirs):
        for (a1, b1, a2, b2) in ((2, 1, 2, 1), (3, 1, 1, 1), (1, 1, 4, 2)):
            for off in ((0, 0), (1, 0), (3, 2), (0, 7), (10, 10), (-5, 3)):
                cases.append((_int_rect(0, 0, *d1, a1, b1), _int_rect(off[0], off[1], *d2, a2, b2)))
    # structured: identical, nested, sharing an edge from outside, touching at one corner, collinear edges with partial overlap
    sq = _int_rect(0, 0, 1, 0, 4, 2)
    cases += [(sq, sq), (sq, _int_rect(0, 0, 1, 0, 2, 1)), (sq, _int_rect(8, 0, 1, 0, 4, 2)), (sq, _int_rect(8, 4, 1, 0, 4, 2)),
              (sq, _int_rect(3, 0, 1, 0, 4, 2)), (sq, _int_rect(0, 0, 0, 1, 4, 2)), (_int_rect(0, 0, 3, 4, 2, 1), _int_rect(0, 0, 3, 4, 2, 1)),
              (_int_rect(0, 0, 3, 4, 2, 1), _int_rect(6, 8, 3, 4, 1, 1)), (_int_rect(0, 0, 1, 1, 3, 3), _int_rect(6, 0, 1, 1, 3, 3))]
    return cases


def _to_np(poly):
    return np.array([[float(x), float(y)] for x, y in poly], dtype=np.float64)


def test_oracle_clip_equals_exact_rational_clip():
    cases = _exact_cases()
    assert len(cases) >= 300
    n_pos = 0
    for p, q in cases:
        exact = _exact_clip_area(p, q)
        mn = float(min(_exact_area(p), _exact_area(q)))
        got = orc.convex_intersection_area(_to_np(p), _to_np(q))
        assert abs(got - float(exact)) <= 1e-12 * mn, (p, q, got, float(exact))
        n_pos += exact > 0
    assert n_pos > 50


def test_device_clip_routine_equals_exact_rational_clip(tmp_path):
    """csrc/mpp_clip.cuh (plain C++ too) in float32 and float64 against the exact areas, in the frame the kernel uses (thinner
    rectangle as the axis-aligned box)."""
    src = tmp_path / "clip_exact.cpp"
    src.write_text('#include <cstdio>\n#include "%s"\n'
                   'extern "C" double area_f64(const double *qx, const double *qy, double hl, double hw) { return mpp_clip::quad_box_area<double>(qx, qy, hl, hw); }\n'
                   'extern "C" double area_f32(const double *qx, const double *qy, double hl, double hw) { float x[4], y[4]; for (int k = 0; k < 4; ++k) { x[k] = (float)qx[k]; y[k] = (float)qy[k]; }\n'
                   '  return (double)mpp_clip::quad_box_area<float>(x, y, (float)hl, (float)hw); }\n' % os.path.join(ROOT, "mpp_cnn_rs_object_detection_b200", "csrc", "mpp_clip.cuh"))
    so = str(tmp_path / "clip_exact.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, str(src)], check=True)
    lib = ctypes.CDLL(so)
    for f in (lib.area_f64, lib.area_f32):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.POINTER(ctypes.c_double)] * 2 + [ctypes.c_double] * 2
    worst64 = worst32 = 0.0
    for p, q in _exact_cases():
        exact = float(_exact_clip_area(p, q))
        A, B = _to_np(p), _to_np(q)
        def sides(P):
            return np.linalg.norm(P[0] - P[3]) / 2, np.linalg.norm(P[0] - P[1]) / 2  # half length (first axis), half width
        if min(sides(B)) < min(sides(A)):
            A, B = B, A
        hl, hw = sides(A)
        c = A.mean(0)
        e0 = (A[0] - A[3]) / (2 * hl)   # unit vector of the length axis
        e1 = (A[0] - A[1]) / (2 * hw)
        loc = np.stack([(B - c) @ e0, (B - c) @ e1], axis=1)
        if 0.5 * np.sum(loc[:, 0] * np.roll(loc[:, 1], -1) - np.roll(loc[:, 0], -1) * loc[:, 1]) > 0:
            loc = loc[::-1].copy()   # the routine takes the quad clockwise
        qx = (ctypes.c_double * 4)(*loc[:, 0])
        qy = (ctypes.c_double * 4)(*loc[:, 1])
        mn = min(4 * hl * hw, orc.poly_area(B))
        worst64 = max(worst64, abs(lib.area_f64(qx, qy, hl, hw) - exact) / mn)
        worst32 = max(worst32, abs(lib.area_f32(qx, qy, hl, hw) - exact) / mn)
    print(f"\nworst |area - exact| / min area: float64 {worst64:.2e}, float32 {worst32:.2e}")
    assert worst64 < 1e-12 and worst32 < 1e-5


def test_oracle_overlap_against_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(17)
    n, worst, n_pos = 10000, 0.0, 0
    for k in range(n):
        a = orc.ORect(0, 0, float(rng.uniform(2, 31)), float(rng.uniform(0.15, 1.0)),
                      float(rng.integers(0, 32) * np.pi / 32 if k % 2 else rng.uniform(0, np.pi)))
        r = int(rng.integers(1, 20))
        b = orc.ORect(int(rng.integers(-r, r + 1)), int(rng.integers(-r, r + 1)), float(rng.uniform(2, 31)), float(rng.uniform(0.15, 1.0)),
                      float(rng.integers(0, 32) * np.pi / 32 if k % 3 == 0 else rng.uniform(0, np.pi)))
        pa, pb = orc.rect_corners(*a.as_tuple()), orc.rect_corners(*b.as_tuple())
        mn = min(orc.poly_area(pa), orc.poly_area(pb))
        want = orc.convex_intersection_area(pa, pb)
        area_cv, _ = cv2.intersectConvexConvex(pa.astype(np.float32), pb.astype(np.float32))
        worst = max(worst, abs(want - float(area_cv)) / mn)
        got = orc.OracleScene.overlap_energy(a, b)  # the C restatement when built, else the same Python routine
        assert abs(got - want / (mn + 1e-6)) < 1e-9
        n_pos += want > 0
    print(f"\n{n} pairs ({n_pos} intersecting): worst |oracle - cv2.intersectConvexConvex| / min area = {worst:.2e}")
    assert n_pos > 2000 and worst <= 3e-5
