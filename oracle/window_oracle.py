"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the WINDOW-RESTRICTED RJMCMC chain (the production sampler's chain).

The production kernels (k_windows_dataflow / k_sweep2, mpp_run_windows) do not run the reference's global chain: each
32x32-px window of a randomly shifted grid runs the reference's kernel mixture *restricted to the window*.  This module
re-derives, proposal by proposal and from first principles, what every such step must compute:

* the kernel drawn (rjmcmc.py:88 with make_kernels.py:76-86 probabilities; births only in an empty window),
* the perturbation (base_kernels.py:31-122, transform_kernels.py:17-225, shape_samplers.py:79-150) given the random words,
* the forward / backward densities of the *restricted* kernels (derivation below),
* the Delta-energy, with the reference's own algorithm (OracleState.delta == energy_graph.py:139-225: recompute the
  3x3-cell blocks before and after; nothing of the device's top-2 partner machinery is used here),
* the accept test (rjmcmc.py:105-113),

and compares each of them with the record the device wrote (mpp_window_trace, include/mpp_b200.h).  All arithmetic here is
float64 on the reference's normalised maps; the device computes in float32 with fast intrinsics, so continuous quantities
are compared within stated tolerances and discrete draws through their CDF interval.

Restricted kernels.  Let Lambda = intensity (base_kernels.py:49), |I| = H*W, |w| = pixels of the window, M_w / M = the
window's / the image's detection mass, n_w = objects in the window.  The reference's global birth proposes a point with
density d(u) relative to the uniform law on the image (d = 1 uniform, shape_samplers.py:143-150; d = det_n * prod(marks_n) *
|I| * 32^3 data-driven, :103-108) and reports forward = p * d / Lambda, backward = p / (n + 1) (base_kernels.py:55-64).
Restricting the birth to the window multiplies the density by |I|/|w| (uniform) or M/M_w (data-driven):
    forward_w = p * d / Lambda_w,   Lambda_w = Lambda * |w| / |I|   (uniform),   Lambda * M_w / M   (data-driven)
    backward_w = p / (n_w + 1)
and symmetrically for deaths (base_kernels.py:94-115).  Moves pick one of the n_w window objects and must end inside the
window (otherwise the proposal is rejected: the restriction is symmetric), so p / n_w cancels; Gaussian kernels are
symmetric (transform_kernels.py:42-58,146-159), the data-driven ones keep det_n(end) / local mass(start) against
det_n(start) / local mass(end) (:94-116) and marks_n(new class) against marks_n(old class) (:205-225).  In an empty window
only births are proposed (probabilities p0/(p0+p2), p2/(p0+p2)); the mixture is state dependent and the Green ratio uses
the probability of the kernel *in the state it is proposed from*.  Each restricted kernel pair is reversible with respect
to the same target as the reference's, so the chain of window visits leaves it invariant.

Only tests/ and tools/ may import this module.  Parity status: pinned transitively -- OracleScene / OracleState are pinned
against reference-generated golden vectors (tests/test_oracle_golden.py); the restriction itself is a derivation, checked
for stationarity against the reference-semantics chain in tests/test_gpu_window_stats.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import mpp_oracle as orc

EPS = orc.EPS
F32 = np.float32
W_EVALUATED, W_ACCEPT, W_IDENTITY, W_HAS_ADD, W_HAS_REM, W_LEFT, W_FULL = 2, 4, 8, 16, 32, 64, 128
_M64 = (1 << 64) - 1
_M32 = 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ random words
def philox4x32_10(seed: int, c0: int, c1: int, c2: int, c3: int) -> Tuple[int, int, int, int]:
    """Philox4x32-10 (Salmon et al., SC'11), key = the 64-bit seed, counter = (c0, c1, c2, c3)."""
    k0, k1 = seed & _M32, (seed >> 32) & _M32
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0 = (k0 + 0x9E3779B9) & _M32
        k1 = (k1 + 0xBB67AE85) & _M32
    return c0, c1, c2, c3


def proposal_words(seed: int, wi: int, wj: int, sweep: int, it: int) -> List[int]:
    """The eight random words of proposal `it` of the visit of window (wi, wj) in `sweep`."""
    c1 = (wi * 65536 + wj) & _M32
    c2 = sweep & _M32
    c3 = ((((sweep >> 32) & _M32) << 20) ^ it ^ 0x77000000) & _M32
    return list(philox4x32_10(seed, 0, c1, c2, c3)) + list(philox4x32_10(seed, 1, c1, c2, c3))


def u01f(b: int) -> F32:
    """24-bit uniform in (0,1), float32 arithmetic."""
    return (F32(b >> 8) + F32(0.5)) * F32(1.0 / 16777216.0)


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def grid_offset(seed: int, sweep: int) -> Tuple[int, int]:
    h = splitmix64((seed & _M64) ^ splitmix64(sweep & _M64))
    return h & 31, (h >> 5) & 31


_STEP = (1.0, 1.0 / 32.0, math.pi / 32.0)
_VMAX = (32.0, 1.0, math.pi)
_EDGES32 = [np.array([F32(k * s) for k in range(32)], dtype=F32) for s in _STEP]


def class_f32(i: int, v) -> int:
    """ValueMapping.value_to_class (mappings.py:45-61) on a float32 value against the float32 bin edges."""
    return int(np.clip(np.searchsorted(_EDGES32[i], F32(v), side="right") - 1, 0, 31))


# ------------------------------------------------------------------------------------------------ objects / scene
class WRect(orc.ORect):
    """ORect that carries the classes of its marks (decided once, where the mark was drawn) next to the values."""
    __slots__ = ("cls",)

    def __init__(self, x, y, size, ratio, angle, uid=-1, cls=None):
        super().__init__(x, y, size, ratio, angle, uid)
        self.cls = tuple(int(c) for c in cls) if cls is not None else tuple(class_f32(i, v) for i, v in enumerate((size, ratio, angle)))


class WindowScene(orc.OracleScene):
    """OracleScene whose mark lookups use the classes carried by the objects (a float32 value sitting exactly on a bin edge
    must not change class when it is widened to float64)."""

    def mark_classes(self, u):
        return u.cls

    def single_mark_energy(self, i, u):
        return self.mark_energy_maps[i][u.x, u.y, u.cls[i]]

    def data_density(self, u) -> float:
        """RectangleSampler.get_point_density shape_samplers.py:103-108."""
        probs = [float(self.marks_n[i][u.x, u.y, u.cls[i]]) for i in range(3)]
        return float(self.detn[u.x, u.y]) * float(np.prod(probs)) * float(self.norm_constant)


def unpack_cls(p: int) -> Tuple[int, int, int]:
    return p & 0xff, (p >> 8) & 0xff, (p >> 16) & 0xff


class Mismatch(AssertionError):
    pass


class WindowOracle:
    """Replays a device trace of the window sampler.  `objects`: iterable of (x, y, size, ratio, angle, uid[, cls])."""

    def __init__(self, scene: WindowScene, combinator, objects, intensity: float, seed: int, per_visit: int,
                 t0: float, alpha_t: float = 1.0, t_target: float = 0.0, p_kernel: Optional[Sequence[float]] = None,
                 trl_sigma: float = orc.GAUSS_TRANSLATION_SIGMA, max_delta: int = orc.DATA_TRANSLATION_MAX_DELTA,
                 trf_sigma: float = orc.GAUSS_TRANSFORM_SIGMA, de_rtol: float = 1e-5, de_atol: float = 2e-5, lr_atol: float = 5e-4,
                 overlap_gain: float = 1.0):
        """overlap_gain: bound of |d energy / d overlap value| of the combinator (0.5 * 0.706 hierarchical, 0.5 * |w_overlap|
        logistic).  It only enters the Delta-energy tolerance of proposals that involve a SLIVER (a rectangle with a side under
        0.2 px, which uniform births draw: ratio ~ U(0,1)): the float32 intersection area is accurate to 5e-6 of the smaller
        area down to 0.1-px half-sides (tools/clip_check.cu) and degrades as 1/half-side below that."""
        self.scene, self.comb = scene, combinator
        self.H, self.W = scene.shape
        self.nx, self.ny = (self.H + 31) // 32, (self.W + 31) // 32
        self.state = orc.OracleState(scene)
        self.by_uid: Dict[int, WRect] = {}
        for o in objects:
            r = WRect(o[0], o[1], float(F32(o[2])), float(F32(o[3])), float(F32(o[4])), uid=int(o[5]), cls=o[6] if len(o) > 6 else None)
            self.state.add(r)
            self.by_uid[r.uid] = r
        self.p = np.asarray(orc.kernel_probabilities() if p_kernel is None else p_kernel, dtype=np.float64)
        self.pf = self.p.astype(F32)
        self.pk_e0 = float(self.p[0] / (self.p[0] + self.p[2]))
        self.pk_e2 = float(self.p[2] / (self.p[0] + self.p[2]))
        self.intensity = float(intensity)
        self.seed, self.pv = int(seed), int(per_visit)
        self.t0, self.alpha_t, self.t_target = float(t0), float(alpha_t), float(t_target)
        self.visit_alpha = F32(self.alpha_t ** (1.0 / self.pv)) if 0.0 < self.alpha_t < 1.0 else F32(1.0)
        self.trl_sigma, self.md = float(trl_sigma), int(max_delta)
        self.trf_sigma = [trf_sigma * v for v in _VMAX]
        self.de_rtol, self.de_atol, self.lr_atol, self.overlap_gain = de_rtol, de_atol, lr_atol, float(overlap_gain)
        self.det64 = scene.det.astype(np.float64)
        self.rowcum = np.concatenate([np.zeros((self.H, 1)), np.cumsum(self.det64, axis=1)], axis=1)
        self.total_mass = float(self.det64.sum())
        self.stats = dict(proposals=0, evaluated=0, accepted=0, identity=0, left_window=0, borderline=0, empty_window=0, slivers=0,
                          max_de_err=0.0, max_lr_err=0.0, per_kernel=[0] * 8, accepted_per_kernel=[0] * 8,
                          evaluated_empty=[0] * 8, evaluated_occupied=[0] * 8, accepted_empty=[0] * 8, accepted_occupied=[0] * 8)

    # ---- helpers
    def pk(self, k: int, n: int) -> float:
        if n > 0:
            return float(self.p[k])
        return self.pk_e0 if k == 0 else (self.pk_e2 if k == 2 else 0.0)

    def _fail(self, where, msg):
        raise Mismatch(f"{where}: {msg}")

    @staticmethod
    def _cdf_consistent(weights: np.ndarray, picked: int, u: float, rel_tol: float = 2e-5) -> bool:
        """Inverse-CDF draw: `picked` must be the first positive-weight index whose inclusive prefix exceeds u * total, up to
        float32 rounding of the prefix sums (the device accumulates in float32)."""
        w = np.asarray(weights, dtype=np.float64)
        tot = float(w.sum())
        if not tot > 0 or not w[picked] > 0:
            return False
        cum = np.cumsum(w)
        t = u * tot
        lo = cum[picked] - w[picked]
        tol = rel_tol * tot
        return lo - tol <= t <= cum[picked] + tol

    def _box_muller(self, a: int, b: int) -> Tuple[float, float]:
        u1, u2 = float(u01f(a)), float(u01f(b))
        r = math.sqrt(-2.0 * math.log(u1))
        return r * math.cos(2.0 * math.pi * u2), r * math.sin(2.0 * math.pi * u2)

    def local_mass(self, x: int, y: int) -> Tuple[int, int, int, int, np.ndarray]:
        md = self.md
        X0, X1, Y0, Y1 = max(0, x - md), min(x + md + 1, self.H), max(0, y - md), min(y + md + 1, self.W)
        return X0, X1, Y0, Y1, self.rowcum[X0:X1, Y1] - self.rowcum[X0:X1, Y0]

    # ---- the chain
    def replay(self, trace: np.ndarray, sweep0: int = 0, check_words: bool = True):
        """trace: WINDOW_TRACE_DTYPE array [n_sweeps, nx + 2, ny + 2, per_visit] (Engine.trace_windows)."""
        n_sweeps = trace.shape[0]
        temp = self.t0
        for s in range(n_sweeps):
            sweep = sweep0 + s
            ox, oy = grid_offset(self.seed, sweep)
            nwx, nwy = (self.H + ox + 31) // 32, (self.W + oy + 31) // 32
            for col in range(9):
                ci, cj = col // 3, col % 3
                for wi in range(ci, nwx, 3):
                    for wj in range(cj, nwy, 3):
                        self.visit(trace[s, wi, wj], sweep, wi, wj, ox, oy, float(F32(temp)), check_words)
            if temp > self.t_target:
                temp *= self.alpha_t
        return self.stats

    def visit_temperature(self, temp: float, it: int) -> float:
        if float(self.visit_alpha) == 1.0:
            return temp
        return max(temp * float(self.visit_alpha) ** it, min(temp, self.t_target))

    def visit(self, recs: np.ndarray, sweep: int, wi: int, wj: int, ox: int, oy: int, temp: float, check_words: bool):
        H, W, pv = self.H, self.W, self.pv
        x0, x1 = max(32 * wi - ox, 0), min(32 * wi - ox + 32, H)
        y0, y1 = max(32 * wj - oy, 0), min(32 * wj - oy + 32, W)
        wx, wy = x1 - x0, y1 - y0
        near = [o for o in self.state if x0 - 64 <= o.x < x1 + 64 and y0 - 64 <= o.y < y1 + 64]
        slots: List[Optional[WRect]] = sorted(near, key=lambda o: (o.x * 16384 + o.y, o.uid))
        in_win = lambda o: o is not None and x0 <= o.x < x1 and y0 <= o.y < y1  # noqa: E731
        uid_base = 0x80000000 | (((sweep * (self.nx + 2) * (self.ny + 2) + wi * (self.ny + 2) + wj) * 128) & 0x7fffffff)
        row_mass = self.rowcum[x0:x1, y1] - self.rowcum[x0:x1, y0]
        win_mass = float(row_mass.sum())
        lam_unif = self.intensity * (wx * wy) / float(H * W)
        lam_data = self.intensity * win_mass / self.total_mass
        st = self.stats
        for it in range(pv):
            rec = recs[it]
            where = f"sweep {sweep} window ({wi},{wj}) proposal {it}"
            flags = int(rec["flags"])
            if not flags & 1:
                self._fail(where, "no trace record")
            q = proposal_words(self.seed, wi, wj, sweep, it)
            if check_words and (int(rec["q"][0]), int(rec["q"][1]), int(rec["q"][2])) != (q[0], q[1], q[7]):
                self._fail(where, "random words differ from Philox4x32-10")
            win_objs = [k for k, o in enumerate(slots) if in_win(o)]
            nc = len(win_objs)
            # ---- kernel choice
            uk = u01f(q[0])
            if nc > 0:
                acc, kernel = F32(0), 7
                for k in range(7):
                    acc = F32(acc + self.pf[k])
                    if uk < acc:
                        kernel = k
                        break
            else:
                kernel = 0 if uk < F32(self.pk_e0) else 2
                st["empty_window"] += 1
            if kernel != ((flags >> 8) & 0xff):
                self._fail(where, f"kernel {(flags >> 8) & 0xff}, oracle {kernel} (n_w = {nc})")
            if min(nc, 255) != ((flags >> 16) & 0xff):
                self._fail(where, f"n_w {(flags >> 16) & 0xff}, oracle {nc}")
            st["proposals"] += 1
            st["per_kernel"][kernel] += 1
            T = self.visit_temperature(temp, it)
            if abs(float(rec["temperature"]) - T) > 2e-5 * T:
                self._fail(where, f"temperature {float(rec['temperature'])}, oracle {T}")
            # ---- the perturbation
            rem: Optional[WRect] = None
            r_slot = -1
            if kernel not in (0, 2):
                r_slot = win_objs[min(nc - 1, int(u01f(q[1]) * F32(nc)))]
                rem = slots[r_slot]
                if not flags & W_HAS_REM or int(rec["rem_uid"]) != rem.uid:
                    self._fail(where, f"removed uid {int(rec['rem_uid'])}, oracle picks {rem.uid}")
            add: Optional[WRect] = None
            log_ratio = 0.0
            expect = "evaluate"  # | 'left' | 'identity'
            dev_add = None
            if flags & W_HAS_ADD:
                dev_add = (int(rec["add_x"]), int(rec["add_y"]), F32(rec["add_size"]), F32(rec["add_ratio"]), F32(rec["add_angle"]),
                           unpack_cls(int(rec["add_cls"])))
            if kernel == 0:  # uniform birth in the window (shape_samplers.py:136-141 restricted to the window)
                ax = x0 + min(wx - 1, int(u01f(q[2]) * F32(wx)))
                ay = y0 + min(wy - 1, int(u01f(q[3]) * F32(wy)))
                vals = (u01f(q[4]) * F32(32.0), u01f(q[5]), u01f(q[6]) * F32(3.14159265358979))
                cls = tuple(class_f32(i, v) for i, v in enumerate(vals))
                if dev_add is None or dev_add[:2] != (ax, ay) or any(dev_add[2 + i] != vals[i] for i in range(3)) or dev_add[5] != cls:
                    self._fail(where, f"uniform birth {dev_add}, oracle {(ax, ay, vals, cls)}")
                add = WRect(ax, ay, *[float(v) for v in vals], uid=uid_base + it, cls=cls)
                log_ratio = math.log(self.pk(1, nc + 1) / (nc + 1) + EPS) - math.log(self.pk(0, nc) / lam_unif + EPS)
            elif kernel == 2:  # data-driven birth (shape_samplers.py:90-98, sampler2d.py:39-46 restricted to the window)
                if not win_mass > 0:
                    expect = "left"
                else:
                    if dev_add is None:
                        self._fail(where, "data-driven birth without an added object")
                    ax, ay, cls = dev_add[0], dev_add[1], dev_add[5]
                    if not (x0 <= ax < x1 and y0 <= ay < y1):
                        self._fail(where, "birth outside the window")
                    ok = self._cdf_consistent(row_mass, ax - x0, float(u01f(q[2]))) and \
                        self._cdf_consistent(self.det64[ax, y0:y1], ay - y0, float(u01f(q[3])))
                    for i in range(3):
                        ok = ok and self._cdf_consistent(self.scene.marks[i][ax, ay], cls[i], float(u01f(q[4 + i])))
                        if dev_add[2 + i] != _EDGES32[i][cls[i]]:
                            self._fail(where, f"mark {i} value {dev_add[2 + i]} is not the lower edge of class {cls[i]}")
                    if not ok:
                        self._fail(where, f"data-driven birth {dev_add} inconsistent with its uniforms")
                    add = WRect(ax, ay, *[float(dev_add[2 + i]) for i in range(3)], uid=uid_base + it, cls=cls)
                    fwd = self.pk(2, nc) * self.scene.data_density(add) / lam_data
                    log_ratio = math.log(self.pk(3, nc + 1) / (nc + 1) + EPS) - math.log(fwd + EPS)
            elif kernel == 1:  # uniform death (base_kernels.py:94-115)
                log_ratio = math.log(self.pk(0, nc - 1) / lam_unif + EPS) - math.log(self.pk(1, nc) / nc + EPS)
            elif kernel == 3:  # data-driven death
                if not win_mass > 0:
                    expect = "left"
                else:
                    bwd = self.pk(2, nc - 1) * self.scene.data_density(rem) / lam_data
                    log_ratio = math.log(bwd + EPS) - math.log(self.pk(3, nc) / nc + EPS)
            elif kernel == 4:  # gaussian translation (transform_kernels.py:24-58)
                d0, d1 = self._box_muller(q[2], q[3])
                cand = []
                for v, lim in ((rem.x + d0 * self.trl_sigma, H - 1), (rem.y + d1 * self.trl_sigma, W - 1)):
                    cand.append({min(max(int(v - 2e-4), 0), lim), min(max(int(v + 2e-4), 0), lim)})
                ax, ay = int(rec["add_x"]), int(rec["add_y"])
                if ax not in cand[0] or ay not in cand[1]:
                    self._fail(where, f"gaussian translation to ({ax},{ay}), oracle {cand}")
                if not (x0 <= ax < x1 and y0 <= ay < y1):
                    expect = "left"
                elif (ax, ay) == (rem.x, rem.y):
                    expect = "identity"
                else:
                    add = WRect(ax, ay, rem.size, rem.ratio, rem.angle, uid=uid_base + it, cls=rem.cls)
            elif kernel == 5:  # data-driven translation (transform_kernels.py:77-116)
                X0, X1, Y0, Y1, rs = self.local_mass(rem.x, rem.y)
                tot_s = float(rs.sum())
                if not tot_s > 0:
                    expect = "left"
                else:
                    ax, ay = int(rec["add_x"]), int(rec["add_y"])
                    if not (X0 <= ax < X1 and Y0 <= ay < Y1):
                        self._fail(where, "data-driven translation outside its 17x17 window")
                    if not (self._cdf_consistent(rs, ax - X0, float(u01f(q[2]))) and
                            self._cdf_consistent(self.det64[ax, Y0:Y1], ay - Y0, float(u01f(q[3])))):
                        self._fail(where, f"data-driven translation to ({ax},{ay}) inconsistent with its uniforms")
                    if not (x0 <= ax < x1 and y0 <= ay < y1):
                        expect = "left"
                    elif (ax, ay) == (rem.x, rem.y):
                        expect = "identity"
                    else:
                        add = WRect(ax, ay, rem.size, rem.ratio, rem.angle, uid=uid_base + it, cls=rem.cls)
                        tot_e = float(self.local_mass(ax, ay)[4].sum())
                        fwd, bwd = self.det64[ax, ay] / tot_s, self.det64[rem.x, rem.y] / tot_e
                        log_ratio = math.log(bwd + EPS) - math.log(fwd + EPS)
            else:  # 6 gaussian / 7 data-driven mark transform (transform_kernels.py:128-225)
                pid = min(2, int(u01f(q[2]) * F32(3.0)))
                if pid != (flags >> 24) & 0xff:
                    self._fail(where, f"param id {(flags >> 24) & 0xff}, oracle {pid}")
                old = (rem.size, rem.ratio, rem.angle)[pid]
                vals, cls = [rem.size, rem.ratio, rem.angle], list(rem.cls)
                if kernel == 6:
                    d0, _ = self._box_muller(q[3], q[4])
                    nv = old + d0 * self.trf_sigma[pid]
                    if pid == 2:
                        nv = nv - math.floor(nv / _VMAX[2]) * _VMAX[2]
                    else:
                        nv = min(max(nv, 0.0), _VMAX[pid])
                    if dev_add is None:
                        self._fail(where, "mark transform without an added object")
                    dv = float(dev_add[2 + pid])
                    wrap = pid == 2 and min(abs(dv - nv), abs(abs(dv - nv) - _VMAX[2])) <= 2e-5 * _VMAX[2]
                    if abs(dv - nv) > 2e-5 * _VMAX[pid] and not wrap:
                        self._fail(where, f"gaussian mark transform value {dv}, oracle {nv}")
                    ncls = class_f32(pid, dev_add[2 + pid])
                    if ncls != dev_add[5][pid]:
                        self._fail(where, f"class of {dv}: device {dev_add[5][pid]}, oracle {ncls}")
                    vals[pid], cls[pid] = dv, ncls
                    add = WRect(rem.x, rem.y, *vals, uid=uid_base + it, cls=cls)
                else:
                    dens = self.scene.marks[pid][rem.x, rem.y]
                    if flags & W_IDENTITY:
                        ncls = rem.cls[pid]
                    else:
                        if dev_add is None:
                            self._fail(where, "mark transform without an added object")
                        ncls = dev_add[5][pid]
                    if not self._cdf_consistent(dens, ncls, float(u01f(q[3]))):
                        self._fail(where, f"data-driven mark transform to class {ncls} inconsistent with its uniform")
                    nv = float(_EDGES32[pid][ncls])
                    if ncls == rem.cls[pid] and nv == old:
                        expect = "identity"
                    else:
                        if float(dev_add[2 + pid]) != nv:
                            self._fail(where, f"mark value {dev_add[2 + pid]} is not the lower edge of class {ncls}")
                        vals[pid], cls[pid] = nv, ncls
                        add = WRect(rem.x, rem.y, *vals, uid=uid_base + it, cls=cls)
                        dn = self.scene.marks_n[pid][rem.x, rem.y]
                        log_ratio = math.log(float(dn[rem.cls[pid]]) + EPS) - math.log(float(dn[ncls]) + EPS)
            # ---- outcome classes that end before the energy
            if flags & W_FULL:
                continue  # destination storage cell full: rejected (capacity limit of the index, reported as such)
            if expect == "left":
                if not flags & W_LEFT or flags & W_EVALUATED:
                    self._fail(where, "the move leaves the window (or has zero proposal mass) but was evaluated")
                st["left_window"] += 1
                continue
            if flags & W_LEFT:
                self._fail(where, "rejected as leaving the window, but the oracle's end point is inside")
            if expect == "identity":
                if not (flags & W_IDENTITY and flags & W_ACCEPT):
                    self._fail(where, "identity proposal not accepted as such")
                st["evaluated"] += 1; st["accepted"] += 1; st["identity"] += 1
                st["accepted_per_kernel"][kernel] += 1
                st["evaluated_occupied"][kernel] += 1; st["accepted_occupied"][kernel] += 1
                continue
            if flags & W_IDENTITY or not flags & W_EVALUATED:
                self._fail(where, f"flags {flags:#x}: expected an evaluated proposal")
            if add is not None:
                if dev_add is None or int(rec["add_uid"]) != add.uid:
                    self._fail(where, f"uid of the addition {int(rec['add_uid'])}, oracle {add.uid}")
                if dev_add[:2] != (add.x, add.y) or dev_add[5] != tuple(add.cls):
                    self._fail(where, f"addition {dev_add}, oracle {add.as_tuple()} {add.cls}")
                for i, v in enumerate((add.size, add.ratio, add.angle)):
                    if float(dev_add[2 + i]) != v:
                        self._fail(where, f"mark {i} of the addition {dev_add[2 + i]}, oracle {v}")
            elif flags & W_HAS_ADD:
                self._fail(where, "unexpected addition")
            # ---- Delta-energy (reference algorithm), Green ratio, accept test
            de = float(self.state.delta([rem] if rem is not None else [], [add] if add is not None else [], self.comb))
            de_dev = float(rec["delta_e"])
            err = abs(de - de_dev)
            thin = min([min(orc.rect_length_width(o.size, o.ratio)) / 2.0 for o in (rem, add) if o is not None] + [1.0])
            de_tol = self.de_atol + self.de_rtol * abs(de)
            if thin < 0.1:
                de_tol += self.overlap_gain * 5e-5 * (0.1 / max(thin, 1e-4))
                st["slivers"] += 1
            else:
                st["max_de_err"] = max(st["max_de_err"], err / max(1.0, abs(de)))
            if err > de_tol:
                self._fail(where, f"Delta-energy {de_dev}, oracle {de} (kernel {kernel})")
            lr_err = abs(log_ratio - float(rec["log_ratio"]))
            st["max_lr_err"] = max(st["max_lr_err"], lr_err)
            if lr_err > self.lr_atol + 1e-5 * abs(log_ratio):
                self._fail(where, f"log(bwd/fwd) {float(rec['log_ratio'])}, oracle {log_ratio} (kernel {kernel}, n_w {nc})")
            u = float(u01f(q[7]))
            if abs(float(rec["u_accept"]) - u) > 1e-7:
                self._fail(where, "accept uniform")
            la = -de / T + log_ratio
            logu = math.log(u + EPS)
            dev_acc = bool(flags & W_ACCEPT)
            margin = de_tol / T + self.lr_atol + 1e-5 * abs(log_ratio)
            if abs(la - logu) <= margin:
                st["borderline"] += 1  # inside the float32 noise of the device's log alpha: follow the device
            elif dev_acc != (logu < la):
                self._fail(where, f"decision {dev_acc}, oracle log u = {logu} vs log alpha = {la} (kernel {kernel})")
            st["evaluated"] += 1
            st["evaluated_occupied" if nc > 0 else "evaluated_empty"][kernel] += 1
            if dev_acc:
                st["accepted"] += 1
                st["accepted_per_kernel"][kernel] += 1
                st["accepted_occupied" if nc > 0 else "accepted_empty"][kernel] += 1
                if rem is not None:
                    self.state.remove(rem)
                    del self.by_uid[rem.uid]
                    slots[r_slot] = None
                if add is not None:
                    self.state.add(add)
                    self.by_uid[add.uid] = add
                    if r_slot >= 0:
                        slots[r_slot] = add
                    else:
                        free = [k for k, o in enumerate(slots) if o is None]
                        if free:
                            slots[free[0]] = add
                        else:
                            slots.append(add)

    def configuration(self) -> np.ndarray:
        """[n, 6] (x, y, size, ratio, angle, uid) sorted by uid."""
        rows = sorted(((o.x, o.y, o.size, o.ratio, o.angle, o.uid) for o in self.state), key=lambda r: r[5])
        return np.array(rows, dtype=np.float64).reshape(-1, 6)
