"""Thin Python owner of one device context (one scene / one chain).  PyTorch is used only for device memory and
streams; every computation is a call into libmpp_b200.so through the C ABI of include/mpp_b200.h."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

LEGACY_NAMES = ["PositionEnergy", "ShapeEnergy", "RectangleOverlapEnergy", "ShapeAlignmentEnergy", "AreaPriorEnergy"]
NOCALIB_NAMES = ["PositionEnergy", "SizeEnergy", "RatioEnergy", "AngleEnergy", "OverlapPriorEnergy", "AlignmentPriorEnergy",
                 "AreaPriorEnergy", "RatioPriorEnergy"]


BASE_KERNEL_WEIGHTS = {"bd_weight": 1, "uniform_bd_weight": 1, "data_bd_weight": 2, "ms_weight": 1, "translation_weight": 1,
                       "gaussian_translation_weight": 1, "data_translation_weight": 2, "transformation_weight": 1,
                       "gaussian_transformation_weight": 1, "data_transformation_weight": 2}  # make_kernels.py:13-24


def kernel_probabilities(use_split_merge: bool = False, weights=None) -> np.ndarray:
    """Kernel choice probabilities of make_kernels (rjmcmc_sampler/kernels/make_kernels.py:13-24,76-86,145-166): 8 entries, or
    10 with the optional split / merge kernels."""
    w = BASE_KERNEL_WEIGHTS if weights is None else weights

    def normalize(a):
        a = np.array(a, dtype=np.float64)
        return a / np.linalg.norm(a, ord=1)

    if use_split_merge:
        p_bd, p_ms, p_trl, p_trf = normalize([w["bd_weight"], w["ms_weight"], w["translation_weight"], w["transformation_weight"]])
    else:
        p_bd, p_trl, p_trf = normalize([w["bd_weight"], w["translation_weight"], w["transformation_weight"]])
        p_ms = None
    p_bd_unif, p_bd_data = normalize([w["uniform_bd_weight"], w["data_bd_weight"]])
    p_trl_gaus, p_trl_data = normalize([w["gaussian_translation_weight"], w["data_translation_weight"]])
    p_trf_gaus, p_trf_data = normalize([w["gaussian_transformation_weight"], w["data_transformation_weight"]])
    p = [0.5 * p_bd_unif * p_bd, 0.5 * p_bd_unif * p_bd, 0.5 * p_bd_data * p_bd, 0.5 * p_bd_data * p_bd,
         p_trl * p_trl_gaus, p_trl * p_trl_data, p_trf * p_trf_gaus, p_trf * p_trf_data]
    if use_split_merge:
        p += [p_ms * 0.5, p_ms * 0.5]
    p = np.array(p)
    if abs(1 - np.sum(p)) < 1e-8:
        p = p / np.sum(p)
    return p


@dataclass
class ModelSpec:
    """Host description of the energy model (terms + combinator) sent to mpp_set_model."""
    setup: str = "legacy"  # 'legacy' | 'nocalib' | 'toy'
    pos_threshold: float = 0.0
    remap_coefs: Sequence[float] = (1.0, 1.0, 1.0)
    remap_intercepts: Sequence[float] = (0.0, 0.0, 0.0)
    min_area: float = 0.0
    max_area: float = 1e9
    ratio_prior: bool = False
    target_ratio: float = 0.5
    rewarding: bool = True
    overlap_max_dist: float = 32.0
    align_max_dist: float = 16.0
    combinator: str = "raw"  # 'raw' | 'hierarchical' | 'logistic' | 'manual'
    comb_w: Sequence[float] = field(default_factory=lambda: [0.0] * 8)
    comb_bias: float = 0.0
    comb_threshold: float = 0.0
    # 'toy' setup (the reference tests' toy terms): constant unit energy + distance-indicator pair energy
    toy_unit_value: float = 0.0
    toy_pair_value: float = 1.0
    toy_pair_dist: float = -1.0
    toy_pair_strict: bool = False
    toy_names: Sequence[str] = ("Unit", "Pair")
    marks_are_energies: bool = False  # the mark maps already hold energies (reference-style pre-computed maps)

    @property
    def names(self):
        if self.setup == "toy":
            return list(self.toy_names)
        if self.setup == "legacy":
            return list(LEGACY_NAMES)
        return list(NOCALIB_NAMES if self.ratio_prior else NOCALIB_NAMES[:7])

    def to_c(self) -> _lib.ModelParams:
        p = _lib.ModelParams()
        p.setup = {"legacy": _lib.SETUP_LEGACY, "nocalib": _lib.SETUP_NO_CALIBRATION, "toy": _lib.SETUP_TOY}[self.setup]
        p.toy_unit_value, p.toy_pair_value = float(self.toy_unit_value), float(self.toy_pair_value)
        p.toy_pair_dist, p.toy_pair_strict = float(self.toy_pair_dist), int(self.toy_pair_strict)
        p.marks_are_energies = int(self.marks_are_energies)
        p.combinator = {"raw": _lib.COMB_RAW_SUM, "hierarchical": _lib.COMB_HIERARCHICAL, "logistic": _lib.COMB_LOGISTIC,
                        "manual": _lib.COMB_MANUAL_HIERARCHICAL}[self.combinator]
        p.ratio_prior = int(self.ratio_prior)
        p.rewarding = int(self.rewarding)
        p.pos_threshold = float(self.pos_threshold)
        for i in range(3):
            p.remap_coef[i] = float(self.remap_coefs[i])
            p.remap_intercept[i] = float(self.remap_intercepts[i])
        p.min_area, p.max_area, p.target_ratio = float(self.min_area), float(self.max_area), float(self.target_ratio)
        p.overlap_max_dist, p.align_max_dist = float(self.overlap_max_dist), float(self.align_max_dist)
        w = list(self.comb_w) + [0.0] * 8
        for i in range(8):
            p.comb_w[i] = float(w[i])
        p.comb_bias, p.comb_threshold = float(self.comb_bias), float(self.comb_threshold)
        return p


def pack_classes(classes: np.ndarray) -> np.ndarray:
    c = np.asarray(classes, dtype=np.uint32).reshape(-1, 3)
    return (c[:, 0] | (c[:, 1] << 8) | (c[:, 2] << 16)).astype(np.uint32)


_EDGES = [np.linspace(0.0, v, 33)[:-1] for v in (32.0, 1.0, math.pi)]  # models/shape_net/mappings.py:17


def classes_of_marks(marks: np.ndarray) -> np.ndarray:
    """ValueMapping.value_to_class in float64 on the host (mappings.py:45-61): max{c : v >= edge_c}."""
    m = np.asarray(marks, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(m.shape, dtype=np.int64)
    for i in range(3):
        c = np.searchsorted(_EDGES[i], m[:, i], side="right") - 1
        if np.any(c < 0):
            raise ValueError(f"mark {i} below its mapping's v_min")
        out[:, i] = np.minimum(c, 31)
    return out


def combine_on_device(model: ModelSpec, vectors: np.ndarray, device=None):
    """EnergyCombinationModel.compute on device: vectors [n, T] -> (per-object [n], total)."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("mpp_cnn_rs_object_detection_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    v = np.asarray(vectors, dtype=np.float64).reshape(len(vectors), -1)
    n = len(v)
    full = np.zeros((n, _lib.MAX_TERMS), dtype=np.float64)
    full[:, :v.shape[1]] = v
    dv = torch.as_tensor(full).to(dev)
    per = torch.zeros(max(n, 1), dtype=torch.float64, device=dev)
    tot = torch.zeros(1, dtype=torch.float64, device=dev)
    p = model.to_c()
    _lib.check(lib.mpp_combine(C.byref(p), dv.data_ptr() if n else None, n, per.data_ptr(), tot.data_ptr(),
                               dev.index if dev.index is not None else torch.cuda.current_device(),
                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return per[:n].cpu().numpy(), float(tot.cpu().item())


_EDGES32 = [e.astype(np.float32) for e in _EDGES]


def snap_to_edges(marks: np.ndarray) -> np.ndarray:
    """Marks read back from a float32 device context: a value that is the float32 rounding of a bin lower edge is
    replaced by the exact float64 edge, i.e. by the value the reference's data-driven kernels would hold
    (ValueMapping.class_to_value, mappings.py:63-74), so that value_to_class in float64 gives the same class."""
    m = np.array(marks, dtype=np.float64).reshape(-1, 3)
    for i in range(3):
        v32 = m[:, i].astype(np.float32)
        k = np.clip(np.searchsorted(_EDGES32[i], v32, side="right") - 1, 0, 31)
        hit = _EDGES32[i][k] == v32
        m[hit, i] = _EDGES[i][k[hit]]
    return m


class Engine:
    """One mpp_ctx.  Not thread-safe; bound to one CUDA device and the current torch stream at creation."""

    _POOL: dict = {}      # (device index, H, W, precision) -> idle contexts (device allocations kept)
    POOL_BYTES = 4 << 30  # idle contexts kept per key: up to this many bytes of device allocations, at least 4, at most 1024 contexts
                          # (a batch of 256 tiles of 512^2 re-uses its 256 contexts: creating and destroying one costs ~25 cudaMalloc / cudaFree)

    @staticmethod
    def _pool_size(shape) -> int:
        h, w = shape
        per_ctx = (h // 32 + 1) * (w // 32 + 1) * (32 * 64 + 16) + h * (w + 1) * 8 + 3 * h * w * 4 + h * w  # records, row prefix sums, class sums, NMS scratch
        return int(max(4, min(1024, Engine.POOL_BYTES // max(1, per_ctx))))

    def __init__(self, shape: Tuple[int, int], device: Optional[torch.device] = None, precision: str = "fp32"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("mpp_cnn_rs_object_detection_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.shape = (int(shape[0]), int(shape[1]))
        self.precision = precision
        self.stream = torch.cuda.current_stream(self.device)
        self._key = (self.device.index, self.shape[0], self.shape[1], precision)
        idle = Engine._POOL.get(self._key)
        if idle:
            ctx = idle.pop()
            _lib.check(self.lib.mpp_ctx_reset(ctx, C.c_void_p(self.stream.cuda_stream)))
        else:
            ctx = C.c_void_p()
            _lib.check(self.lib.mpp_ctx_create(C.byref(ctx), self.device.index, self.shape[0], self.shape[1],
                                               _lib.PRECISION_FP64 if precision == "fp64" else _lib.PRECISION_FP32,
                                               C.c_void_p(self.stream.cuda_stream)))
        self.ctx = ctx
        self.model: Optional[ModelSpec] = None
        self._det = None
        self._marks = None
        self.launches = 0  # kernels launched through this engine (bench gpu_launches)

    def close(self):
        """Returns the context to the pool (its device buffers are reused by the next Engine of the same shape)."""
        ctx = getattr(self, "ctx", None)
        if ctx is not None and ctx:
            self.ctx = None
            self._det = self._marks = None
            idle = Engine._POOL.setdefault(self._key, [])
            if len(idle) < Engine._pool_size(self.shape):
                idle.append(ctx)
            else:
                self.lib.mpp_ctx_destroy(ctx)

    @classmethod
    def drain_pool(cls):
        lib = _lib.load()
        for idle in cls._POOL.values():
            while idle:
                lib.mpp_ctx_destroy(idle.pop())

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------ setup
    def _dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def set_maps(self, det, marks, det_sum: Optional[float] = None):
        """det (H,W) f32; marks (3,H,W,32) f32 or a list of three (H,W,32).  Tensors already on the device are used
        in place (no copy).  det_sum: float(np.sum(det)) as numpy computes it, for bit-parity of the densities."""
        if isinstance(marks, (list, tuple)):
            if isinstance(marks[0], torch.Tensor):
                marks = torch.stack([m.to(self.device) for m in marks])
            else:
                marks = np.stack([np.asarray(m, dtype=np.float32) for m in marks])
        if det_sum is None and not isinstance(det, torch.Tensor):
            det_sum = float(np.sum(np.asarray(det, dtype=np.float32)))
        self._det = self._dev(det, torch.float32)
        self._marks = self._dev(marks, torch.float32)
        assert tuple(self._det.shape) == self.shape, "detection map shape"
        assert tuple(self._marks.shape) == (3,) + self.shape + (32,), "mark maps shape"
        _lib.check(self.lib.mpp_set_maps(self.ctx, self._det.data_ptr(), self._marks.data_ptr(),
                                         float(det_sum) if det_sum is not None else -1.0))
        self.launches += 2

    def set_maps_band(self, det_band, marks_band, row0: int, det_sum_scene: float):
        """Band-local maps of a split scene (mpp_set_maps_band): det_band (rows, W), marks_band (3, rows, W, 32) = the scene rows
        [row0, row0 + rows); det_sum_scene = sum of the whole detection map."""
        self._det = self._dev(det_band, torch.float32)
        self._marks = self._dev(marks_band, torch.float32)
        rows = int(self._det.shape[0])
        assert tuple(self._det.shape) == (rows, self.shape[1]) and tuple(self._marks.shape) == (3, rows, self.shape[1], 32)
        _lib.check(self.lib.mpp_set_maps_band(self.ctx, self._det.data_ptr(), self._marks.data_ptr(), int(row0), rows, float(det_sum_scene)))
        self.launches += 5

    def set_model(self, model: ModelSpec):
        self.model = model
        p = model.to_c()
        _lib.check(self.lib.mpp_set_model(self.ctx, C.byref(p)))

    def set_kernels(self, intensity: float, p_kernel: Optional[np.ndarray] = None, translation_sigma: float = 2.0,
                    max_delta: int = 8, transform_sigma: float = 0.1):
        p = _lib.KernelParams()
        pk = kernel_probabilities() if p_kernel is None else np.asarray(p_kernel, dtype=np.float64)
        for i in range(8):
            p.p_kernel[i] = float(pk[i])
        p.intensity = float(intensity)
        p.gauss_translation_sigma = float(translation_sigma)
        p.data_translation_max_delta = int(max_delta)
        p.gauss_transform_sigma = float(transform_sigma)
        _lib.check(self.lib.mpp_set_kernels(self.ctx, C.byref(p)))

    # ------------------------------------------------------------------------------------------ objects
    def add_objects(self, xy, marks, classes=None, uid=None) -> np.ndarray:
        xy = np.ascontiguousarray(np.asarray(xy, dtype=np.int32).reshape(-1, 2))
        marks = np.ascontiguousarray(np.asarray(marks, dtype=np.float64).reshape(-1, 3))
        n = len(xy)
        if n == 0:
            return np.zeros((0,), dtype=np.uint32)
        if classes is None:
            classes = classes_of_marks(marks)
        d_xy, d_m = self._dev(xy, torch.int32), self._dev(marks, torch.float64)
        d_c = self._dev(pack_classes(classes).astype(np.int64), torch.int64).to(torch.int32)  # bit pattern of u32
        d_u = None if uid is None else self._dev(np.asarray(uid, dtype=np.int64), torch.int64).to(torch.int32)
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.mpp_add_objects(self.ctx, d_xy.data_ptr(), d_m.data_ptr(), d_c.data_ptr(),
                                            None if d_u is None else d_u.data_ptr(), n, out.data_ptr()))
        self.launches += 1
        return out.cpu().numpy().view(np.uint32)

    def remove_objects(self, handles):
        h = np.ascontiguousarray(np.asarray(handles, dtype=np.uint32).reshape(-1))
        if len(h) == 0:
            return
        d = self._dev(h.view(np.int32), torch.int32)
        _lib.check(self.lib.mpp_remove_objects(self.ctx, d.data_ptr(), len(h)))
        self.launches += 1

    def clear(self):
        _lib.check(self.lib.mpp_clear_objects(self.ctx))

    def __len__(self):
        n = C.c_int()
        _lib.check(self.lib.mpp_num_objects(self.ctx, C.byref(n)))
        return n.value

    def read_objects(self):
        """(handles u32 [N], xy i32 [N,2], marks f64 [N,3], uid u32 [N]) in cell-major order."""
        n = len(self)
        cap = max(n, 1)
        h = torch.empty(cap, dtype=torch.int32, device=self.device)
        xy = torch.empty((cap, 2), dtype=torch.int32, device=self.device)
        m = torch.empty((cap, 3), dtype=torch.float64, device=self.device)
        u = torch.empty(cap, dtype=torch.int32, device=self.device)
        cnt = C.c_int()
        _lib.check(self.lib.mpp_read_objects(self.ctx, cap, h.data_ptr(), xy.data_ptr(), m.data_ptr(), u.data_ptr(), C.byref(cnt)))
        self.launches += 2
        n = min(cnt.value, cap)
        marks = m[:n].cpu().numpy()
        if self.precision != "fp64":
            marks = snap_to_edges(marks)
        return (h[:n].cpu().numpy().view(np.uint32), xy[:n].cpu().numpy(), marks, u[:n].cpu().numpy().view(np.uint32))

    # ------------------------------------------------------------------------------------------ energies
    def energy_vectors(self, handles):
        """(vectors f64 [N,T], combined f64 [N], raw_total, combined_total) for the objects named by `handles`."""
        h = np.ascontiguousarray(np.asarray(handles, dtype=np.uint32).reshape(-1))
        n = len(h)
        t = len(self.model.names)
        if n == 0:
            return np.zeros((0, t)), np.zeros((0,)), 0.0, 0.0
        d = self._dev(h.view(np.int32), torch.int32)
        vec = torch.empty((n, _lib.MAX_TERMS), dtype=torch.float64, device=self.device)
        comb = torch.empty(n, dtype=torch.float64, device=self.device)
        tot = torch.empty(2, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mpp_energy_vectors(self.ctx, d.data_ptr(), n, vec.data_ptr(), comb.data_ptr(), tot.data_ptr()))
        self.launches += 2
        tot = tot.cpu().numpy()
        return vec.cpu().numpy()[:, :t], comb.cpu().numpy(), float(tot[0]), float(tot[1])

    def delta_batch(self, proposals: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(proposals, dtype=_lib.PROPOSAL_DTYPE)
        m = len(p)
        if m == 0:
            return np.zeros((0,))
        d = torch.as_tensor(p.view(np.uint8).reshape(m, -1)).to(self.device)
        out = torch.empty(m, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mpp_delta_batch(self.ctx, d.data_ptr(), m, out.data_ptr()))
        self.launches += 1
        return out.cpu().numpy()

    def replay(self, proposals: np.ndarray, t0: float, alpha_t: float, t_target: float = 0.0) -> np.ndarray:
        p = np.ascontiguousarray(proposals, dtype=_lib.PROPOSAL_DTYPE)
        m = len(p)
        d = torch.as_tensor(p.view(np.uint8).reshape(m, -1)).to(self.device)
        out = torch.zeros((m, _lib.STEP_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mpp_replay(self.ctx, d.data_ptr(), m, float(t0), float(alpha_t), float(t_target), out.data_ptr()))
        self.launches += 1
        return out.cpu().numpy().view(_lib.STEP_RESULT_DTYPE).reshape(m)

    def run_sweeps(self, n_sweeps: int, proposals_per_visit: int = 8, stride: int = 4, t0: float = 1.0, alpha_t: float = 1.0,
                   t_target: float = 0.0, seed: int = 0, sweep_offset: int = 0, read_counters: bool = True):
        cnt = (C.c_ulonglong * 8)()
        _lib.check(self.lib.mpp_run_sweeps(self.ctx, int(n_sweeps), int(proposals_per_visit), int(stride), float(t0),
                                           float(alpha_t), float(t_target), int(seed), int(sweep_offset),
                                           cnt if read_counters else None))
        self.launches += int(n_sweeps) * stride * stride
        return [int(v) for v in cnt] if read_counters else None

    def run_windows(self, n_sweeps: int, proposals_per_visit: int = 96, n_warps: int = 8, t0: float = 1.0, alpha_t: float = 1.0,
                    t_target: float = 0.0, seed: int = 0, sweep_offset: int = 0, read_counters: bool = True, debug: bool = False,
                    schedule: str = "dataflow"):
        """Production parallel sampler (mpp_run_windows): shifted 32-px windows, shared-memory resident visits, speculative
        evaluation by `n_warps` warps (1, 2, 4, 8: one proposal per warp; 0: one warp per window, one proposal per lane).  Returns [proposals, accepted, births, deaths, evaluated, 0, 0, 0] (+ the largest
        |fast - brute-force| Delta-energy difference when debug=True)."""
        cnt = (C.c_ulonglong * 8)()
        dbg = torch.zeros(1, dtype=torch.float32, device=self.device) if debug else None
        sched = {"colours": 0, "dataflow": 1}[schedule]
        _lib.check(self.lib.mpp_run_windows(self.ctx, int(n_sweeps), int(proposals_per_visit), int(n_warps), sched, float(t0), float(alpha_t),
                                            float(t_target), int(seed), int(sweep_offset), cnt if read_counters else None,
                                            None if dbg is None else dbg.data_ptr()))
        self.launches += (1 if n_sweeps > 0 else 0) if sched == 1 else int(n_sweeps) * 9
        out = [int(v) for v in cnt] if read_counters else None
        if debug:
            return out, float(dbg.cpu().item())
        return out

    # ------------------------------------------------------------------------------------------ scene split across GPUs
    def split_export(self) -> bytes:
        """CUDA IPC handles of this context's occupancy masks, records and completion grid (mpp_split_export)."""
        buf = C.create_string_buffer(3 * _lib.IPC_HANDLE_BYTES)
        _lib.check(self.lib.mpp_split_export(self.ctx, buf))
        return buf.raw

    def split_attach(self, row_lo: int, row_hi: int, up_handles: Optional[bytes], down_handles: Optional[bytes]):
        """This context samples the band [row_lo, row_hi) of its scene; the neighbour bands live in other processes."""
        _lib.check(self.lib.mpp_split_attach(self.ctx, int(row_lo), int(row_hi), up_handles, down_handles))

    def split_attach_local(self, row_lo: int, row_hi: int, up: Optional["Engine"], down: Optional["Engine"]):
        _lib.check(self.lib.mpp_split_attach_local(self.ctx, int(row_lo), int(row_hi), None if up is None else up.ctx,
                                                   None if down is None else down.ctx))

    def split_detach(self):
        _lib.check(self.lib.mpp_split_detach(self.ctx))

    def window_stats(self) -> dict:
        """Per-kernel tallies of the window sampler since the last call (mpp_window_stats; reads and resets)."""
        raw = (C.c_ulonglong * _lib.WINDOW_STATS)()
        _lib.check(self.lib.mpp_window_stats(self.ctx, raw))
        v = [int(x) for x in raw]
        return {"evaluated_empty": v[0:8], "evaluated_occupied": v[8:16], "accepted_empty": v[16:24], "accepted_occupied": v[24:32],
                "identity_accepted": v[32], "visits": v[33], "visits_empty": v[34]}

    def trace_windows(self, n_sweeps: int, proposals_per_visit: int, sweep_offset: int = 0, **kw):
        """run_windows(debug=True) with the per-proposal trace armed (mpp_set_window_trace).  Returns (counters, max |fast -
        brute-force Delta E|, trace) where trace is a WINDOW_TRACE_DTYPE array of shape [n_sweeps, nx + 2, ny + 2,
        proposals_per_visit] (entries of windows that do not exist in a sweep stay zero)."""
        nx, ny = (self.shape[0] + 31) // 32, (self.shape[1] + 31) // 32
        n = int(n_sweeps) * (nx + 2) * (ny + 2) * int(proposals_per_visit)
        buf = torch.zeros((max(n, 1), _lib.WINDOW_TRACE_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mpp_set_window_trace(self.ctx, buf.data_ptr(), n, int(sweep_offset)))
        try:
            cnt, maxdiff = self.run_windows(n_sweeps, proposals_per_visit, sweep_offset=sweep_offset, debug=True, **kw)
        finally:
            _lib.check(self.lib.mpp_set_window_trace(self.ctx, None, 0, 0))
        tr = buf.cpu().numpy().view(_lib.WINDOW_TRACE_DTYPE).reshape(-1)[:n]
        return cnt, maxdiff, tr.reshape(int(n_sweeps), nx + 2, ny + 2, int(proposals_per_visit))

    def run_window_rows(self, proposals_per_visit: int, n_warps: int, temperature: float, seed: int, sweep_id: int, ci: int,
                        row_lo: int, row_hi: int):
        """One third of a sweep (window rows wi = ci mod 3) restricted to the rows [row_lo, row_hi) (scene split across GPUs)."""
        _lib.check(self.lib.mpp_run_window_rows(self.ctx, int(proposals_per_visit), int(n_warps), float(temperature), int(seed),
                                                int(sweep_id), int(ci), int(row_lo), int(row_hi)))
        self.launches += 3

    def window_grid(self, seed: int, sweep_id: int):
        ox, oy = C.c_int(), C.c_int()
        _lib.check(self.lib.mpp_window_grid(self.ctx, int(seed), int(sweep_id), C.byref(ox), C.byref(oy)))
        return ox.value, oy.value

    def run_chain(self, n_steps: int, t0: float = 1.0, alpha_t: float = 1.0, t_target: float = 0.0, seed: int = 0,
                  step_offset: int = 0, trace: bool = False, read_counters: bool = True):
        """Sequential device chain with the reference's global kernels (RJMCMC.run, rjmcmc.py:83-181)."""
        tr = torch.zeros((max(n_steps, 1), _lib.STEP_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=self.device) if trace else None
        cnt = (C.c_ulonglong * 8)()
        _lib.check(self.lib.mpp_run_chain(self.ctx, int(n_steps), float(t0), float(alpha_t), float(t_target), int(seed),
                                          int(step_offset), None if tr is None else tr.data_ptr(), cnt if read_counters else None))
        self.launches += 2 if n_steps > 0 else 0
        counters = [int(v) for v in cnt] if read_counters else None
        if trace:
            return counters, tr.cpu().numpy().view(_lib.STEP_RESULT_DTYPE).reshape(-1)[:n_steps]
        return counters

    def sample_proposals(self, kernel_ids, seed: int = 0, offset: int = 0) -> np.ndarray:
        """Kernel.sample_perturbation for each entry of kernel_ids (-1: draw the kernel too) against the current state."""
        k = np.ascontiguousarray(np.asarray(kernel_ids, dtype=np.int32).reshape(-1))
        m = len(k)
        if m == 0:
            return np.zeros(0, dtype=_lib.PROPOSAL_DTYPE)
        dk = self._dev(k, torch.int32)
        out = torch.zeros((m, _lib.PROPOSAL_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mpp_sample_proposals(self.ctx, dk.data_ptr(), m, int(seed), int(offset), out.data_ptr()))
        self.launches += 2
        rec = out.cpu().numpy().view(_lib.PROPOSAL_DTYPE).reshape(m).copy()
        if self.precision != "fp64":
            snapped = snap_to_edges(np.stack([rec["add_size"], rec["add_ratio"], rec["add_angle"]], axis=1))
            rec["add_size"], rec["add_ratio"], rec["add_angle"] = snapped[:, 0], snapped[:, 1], snapped[:, 2]
        return rec

    def proposal_probs(self, proposals: np.ndarray) -> np.ndarray:
        """[m,2] forward / backward probabilities (Kernel.forward_probability / backward_probability)."""
        p = np.ascontiguousarray(proposals, dtype=_lib.PROPOSAL_DTYPE)
        m = len(p)
        if m == 0:
            return np.zeros((0, 2))
        d = torch.as_tensor(p.view(np.uint8).reshape(m, -1)).to(self.device)
        out = torch.empty((m, 2), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mpp_proposal_probs(self.ctx, d.data_ptr(), m, out.data_ptr()))
        self.launches += 1
        return out.cpu().numpy()

    def sample_split_merge(self, kind: int, radius: float, shape_sigmas, seed: int = 0, offset: int = 0) -> np.ndarray:
        """SplitKernel (kind 8) / MergeKernel (kind 9) .sample_perturbation on the device: one SPLIT_MERGE_DTYPE record."""
        sig = (C.c_double * 3)(*[float(v) for v in shape_sigmas])
        out = torch.zeros(_lib.SPLIT_MERGE_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.mpp_sample_split_merge(self.ctx, int(kind), float(radius), sig, int(seed), int(offset), out.data_ptr()))
        self.launches += 2
        return out.cpu().numpy().view(_lib.SPLIT_MERGE_DTYPE).reshape(1).copy()

    def split_merge_probs(self, record: np.ndarray, p_split: float, p_merge: float, radius: float, shape_sigmas):
        """(forward, backward) densities of one split / merge perturbation against the current state, on the device."""
        rec = np.ascontiguousarray(record, dtype=_lib.SPLIT_MERGE_DTYPE).reshape(1)
        d = torch.as_tensor(rec.view(np.uint8).reshape(-1)).to(self.device)
        sig = (C.c_double * 3)(*[float(v) for v in shape_sigmas])
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mpp_split_merge_probs(self.ctx, d.data_ptr(), float(p_split), float(p_merge), float(radius), sig, out.data_ptr()))
        self.launches += 2
        v = out.cpu().numpy()
        return float(v[0]), float(v[1])

    def sample_births(self, n: int, seed: int = 0) -> np.ndarray:
        out = torch.empty((n, 5), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.mpp_sample_births(self.ctx, int(n), int(seed), out.data_ptr()))
        self.launches += 1
        return out.cpu().numpy()

    def naive_init(self, detection_threshold: float, nms_distance: float = 6.0) -> int:
        n = C.c_int()
        _lib.check(self.lib.mpp_naive_init(self.ctx, float(detection_threshold), float(nms_distance), C.byref(n)))
        return n.value

    def query_neighbors(self, x: int, y: int, radius: float, euclidean: bool = False, exclude: int = _lib.NO_OBJECT,
                        capacity: int = 1024) -> np.ndarray:
        """Handles of PointsSet.get_potential_neighbors / get_neighbors (point_set.py:111-149)."""
        while True:
            out = torch.empty(max(capacity, 1), dtype=torch.int32, device=self.device)
            n = C.c_int()
            _lib.check(self.lib.mpp_query_neighbors(self.ctx, int(x), int(y), float(radius), int(bool(euclidean)),
                                                    C.c_uint32(int(exclude)), capacity, out.data_ptr(), C.byref(n)))
            self.launches += 1
            if n.value <= capacity:
                return out[:n.value].cpu().numpy().view(np.uint32)
            capacity = n.value

    def pair_values(self, handles_a, handles_b) -> np.ndarray:
        """[n,2] float64 (overlap kind, alignment kind) of the given pairs; NaN where the pair does not exist."""
        a = np.ascontiguousarray(np.asarray(handles_a, dtype=np.uint32).reshape(-1))
        b = np.ascontiguousarray(np.asarray(handles_b, dtype=np.uint32).reshape(-1))
        assert len(a) == len(b)
        if len(a) == 0:
            return np.zeros((0, 2))
        da, db = self._dev(a.view(np.int32), torch.int32), self._dev(b.view(np.int32), torch.int32)
        out = torch.empty((len(a), 2), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mpp_pair_values(self.ctx, da.data_ptr(), db.data_ptr(), len(a), out.data_ptr()))
        self.launches += 1
        return out.cpu().numpy()

    def copy_state_from(self, other: "Engine"):
        _lib.check(self.lib.mpp_copy_state(self.ctx, other.ctx))

    def pack_rows(self, row_lo: int, row_hi: int, capacity: int = 65536) -> torch.Tensor:
        buf = torch.empty((capacity, 8), dtype=torch.float64, device=self.device)
        n = C.c_int()
        _lib.check(self.lib.mpp_pack_rows(self.ctx, int(row_lo), int(row_hi), buf.data_ptr(), capacity, C.byref(n)))
        return buf[:n.value]

    def unpack_rows(self, row_lo: int, row_hi: int, records: torch.Tensor):
        rec = records.to(device=self.device, dtype=torch.float64).contiguous()
        _lib.check(self.lib.mpp_unpack_rows(self.ctx, int(row_lo), int(row_hi), rec.data_ptr() if len(rec) else None, len(rec)))


def run_windows_batch(engines: Sequence[Engine], seeds: Sequence[int], n_sweeps: int, proposals_per_visit: int = 96, n_warps: int = 8,
                      t0: float = 1.0, alpha_t: float = 1.0, t_target: float = 0.0, grid_seed: Optional[int] = None, sweep_offset: int = 0,
                      max_ctas: int = 0, read_counters: bool = True, debug: bool = False):
    """mpp_run_windows_batch: the window sampler over a batch of scenes of equal shape (or over this rank's band of a split
    scene: one engine) in ONE persistent launch on the first engine's stream.  Returns the summed counters
    [proposals, accepted, births, deaths, evaluated, 0, 0, 0] (and the largest |fast - brute-force Delta E| with debug=True)."""
    engines = list(engines)
    n = len(engines)
    lib = engines[0].lib
    ctxs = (C.c_void_p * n)(*[e.ctx for e in engines])
    sd = (C.c_uint64 * n)(*[int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds])
    cnt = (C.c_ulonglong * 8)()
    dbg = torch.zeros(1, dtype=torch.float32, device=engines[0].device) if debug else None
    gs = int(seeds[0] if grid_seed is None else grid_seed) & 0xFFFFFFFFFFFFFFFF
    _lib.check(lib.mpp_run_windows_batch(ctxs, sd, n, gs, int(n_sweeps), int(proposals_per_visit), int(n_warps), float(t0), float(alpha_t),
                                         float(t_target), int(sweep_offset), int(max_ctas), cnt if read_counters else None,
                                         None if dbg is None else dbg.data_ptr()))
    engines[0].launches += 1 if n_sweeps > 0 else 0
    out = [int(v) for v in cnt] if read_counters else None
    if debug:
        return out, float(dbg.cpu().item())
    return out
