"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Makes the *unmodified* reference at /root/reference importable in this container so
that golden vectors can be generated from it (oracle/gen_golden.py).  The reference
needs three third-party modules that are absent from this image:

  * skimage.draw      -- drawing only (base/shapes/base_shapes.py:7), never on the hot path
  * matplotlib        -- plotting only (base/shapes/rectangle.py:5)
  * shapely.geometry  -- Polygon(...).area / .intersection(...).area
                         (models/mpp/energies/prior_energies.py:14-18, :63)

The first two are replaced by inert dummies.  The shapely stub is the only one that
carries arithmetic: a float64 Sutherland-Hodgman clip + shoelace area, which is exact
(up to rounding) for the convex quadrilaterals the reference feeds it.  Because the
real shapely==1.7.1 / geos==3.8.0 (env.yml:184,50) cannot be installed here (no
network) and no reference test pins values at that boundary, PARITY IS UNPINNED at the
shapely boundary; everything else in the golden vectors is the reference's own code.

Stubs are only installed when the real module cannot be imported.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


class _Dummy:
    """Callable / attribute sink used for plotting-only modules."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())


def _dummy_module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr: _Dummy()  # type: ignore[attr-defined]
    mod.__path__ = []  # behave like a package
    return mod


# --------------------------------------------------------------------------- shapely stub
def _signed_area(p: np.ndarray) -> float:
    x, y = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _clip_convex(subject: np.ndarray, clip: np.ndarray) -> np.ndarray:
    """Sutherland-Hodgman: clip convex `subject` by convex `clip` (any orientation)."""
    if _signed_area(clip) < 0:
        clip = clip[::-1]
    out = [tuple(v) for v in subject]
    n = len(clip)
    for i in range(n):
        if not out:
            break
        ax, ay = clip[i]
        bx, by = clip[(i + 1) % n]
        ex, ey = bx - ax, by - ay
        inp, out = out, []
        m = len(inp)
        for k in range(m):
            px, py = inp[k]
            qx, qy = inp[(k + 1) % m]
            sp = ex * (py - ay) - ey * (px - ax)
            sq = ex * (qy - ay) - ey * (qx - ax)
            if sp >= 0:
                out.append((px, py))
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                out.append((px + t * (qx - px), py + t * (qy - py)))
    return np.array(out, dtype=np.float64).reshape(-1, 2)


class StubPolygon:
    """Minimal stand-in for shapely.geometry.Polygon restricted to convex rings."""

    def __init__(self, coords=None):
        self._c = np.zeros((0, 2)) if coords is None else np.asarray(coords, dtype=np.float64).reshape(-1, 2)

    @property
    def area(self) -> float:
        if len(self._c) < 3:
            return 0.0
        return abs(_signed_area(self._c))

    def intersection(self, other: "StubPolygon") -> "StubPolygon":
        # a zero-area ring (size == 0 or ratio == 0 after clipping, transform_kernels.py:138-139) has an empty
        # interior: the intersection has zero area (GEOS treats such rings as invalid; unpinned either way)
        if len(self._c) < 3 or len(other._c) < 3 or self.area == 0.0 or other.area == 0.0:
            return StubPolygon()
        return StubPolygon(_clip_convex(self._c, other._c))


def install(reference_root: str = REFERENCE_ROOT) -> dict:
    """Install stubs for missing modules and put the reference on sys.path.

    Returns a dict module-name -> 'real' | 'stub' so callers can record provenance.
    """
    # import the heavy real deps first: torch introspects matplotlib if it finds a fake one
    import scipy.stats  # noqa: F401
    import sklearn.metrics  # noqa: F401
    import torch  # noqa: F401

    status = {}
    for name in ("skimage", "matplotlib", "shapely"):
        try:
            importlib.import_module(name)
            status[name] = "real"
        except Exception:
            status[name] = "stub"

    if status["skimage"] == "stub":
        sk = _dummy_module("skimage")
        dr = _dummy_module("skimage.draw")
        dr.draw = dr  # reference does `from skimage.draw import draw`
        sk.draw = dr
        sys.modules["skimage"] = sk
        sys.modules["skimage.draw"] = dr
    if status["matplotlib"] == "stub":
        mp = _dummy_module("matplotlib")
        for sub in ("pyplot", "patches", "cm", "colors"):
            m = _dummy_module(f"matplotlib.{sub}")
            setattr(mp, sub, m)
            sys.modules[f"matplotlib.{sub}"] = m
        sys.modules["matplotlib"] = mp
    if status["shapely"] == "stub":
        sh = types.ModuleType("shapely")
        geo = types.ModuleType("shapely.geometry")
        geo.Polygon = StubPolygon
        sh.geometry = geo
        sh.__path__ = []
        sys.modules["shapely"] = sh
        sys.modules["shapely.geometry"] = geo

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return status
