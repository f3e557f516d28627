"""Glue between the reference-style Python objects and one device context (Engine): translation of term constructors
into the device energy model, a cache of device-resident maps, and the host mirror that gives Python objects their
identity (the reference hashes objects by id(), base/shapes/base_shapes.py:16-17; on the device an object is a
(cell, slot) handle)."""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from ..engine import Engine, ModelSpec, classes_of_marks
from .energies import (AreaPriorEnergy, ConstantUnitEnergy, DistanceIndicatorPairEnergy, PairEnergyConstructor, PositionEnergy,
                       RatioPriorEnergy, RectangleOverlapEnergy, ShapeAlignmentEnergy, ShapeEnergy, SingleMarkEnergy,
                       UnitEnergyConstructor)
from .shapes import Point, Rectangle

MAX_DEVICE_RADIUS = 32  # the device grid has 32-px cells (MPP_CELL_SIZE); point_set.py:58 uses max(r_max, 32)


# ---------------------------------------------------------------------------------------------- maps
class DeviceMaps:
    """Detection map (H,W) and mark maps (3,H,W,32) resident on the device."""

    def __init__(self, det: torch.Tensor, marks: torch.Tensor, det_sum: Optional[float]):
        self.det, self.marks, self.det_sum = det, marks, det_sum


_MAPS_CACHE: Dict[Tuple[int, ...], DeviceMaps] = {}
# idle device buffers for uploads (reuse=False): (device index, H, W) -> [(det buffer, marks buffer)].  cudaMalloc /
# cudaFree of a 1.6 GB map set costs far more than the upload itself, so the buffers are recycled.
_UPLOAD_POOL: Dict[Tuple[int, int, int], list] = {}
_UPLOAD_POOL_SIZE = 3


def _upload_buffers(dev: torch.device, h: int, w: int):
    idle = _UPLOAD_POOL.get((dev.index, h, w))
    if idle:
        return idle.pop()
    return (torch.empty((h, w), dtype=torch.float32, device=dev), torch.empty((3, h, w, 32), dtype=torch.float32, device=dev))


def _recycle(key, bufs):
    idle = _UPLOAD_POOL.setdefault(key, [])
    if len(idle) < _UPLOAD_POOL_SIZE:
        idle.append(bufs)


def _evict(key):
    _MAPS_CACHE.pop(key, None)


def _array_key(a):
    """Cache key of one source array: where its data lives, its shape / dtype, and something that changes when it is edited in
    place -- the version counter of a tensor, a strided sample of a numpy array (about 1e4 elements; a cheap fingerprint, not a
    proof: callers that rewrite their maps in place should pass reuse_device_maps=False)."""
    if isinstance(a, torch.Tensor):
        return ("t", a.data_ptr(), tuple(a.shape), str(a.dtype), a._version)
    arr = np.asarray(a)
    flat = arr.reshape(-1)
    step = max(1, flat.size // 8191)
    return ("n", arr.__array_interface__["data"][0], arr.shape, str(arr.dtype), float(np.sum(flat[::step], dtype=np.float64)))


def device_maps(det, marks: Sequence, device=None, reuse: bool = True) -> DeviceMaps:
    """Uploads the maps to the device (once per set of source arrays when reuse=True).  Tensors already on the device are
    used in place; pinned host tensors are copied asynchronously on the current stream."""
    stacked = isinstance(marks, torch.Tensor) and marks.dim() == 4
    sources = [det, marks] if stacked else [det] + list(marks)
    key = tuple(_array_key(a) for a in sources) if reuse else None
    hit = _MAPS_CACHE.get(key) if reuse else None
    if hit is not None:
        return hit
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    det_on_dev = isinstance(det, torch.Tensor) and det.is_cuda
    pooled = None
    if not reuse and not det_on_dev:
        hh, ww = int(det.shape[0]), int(det.shape[1])
        pooled = _upload_buffers(dev, hh, ww)
    if isinstance(det, torch.Tensor):
        det_sum = None
        if pooled is not None:
            d = pooled[0]
            d.copy_(det, non_blocking=True)
        else:
            d = det.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    else:
        det_np = np.ascontiguousarray(det, dtype=np.float32)
        det_sum = float(np.sum(det_np))  # shape_samplers.py:87 normalises with numpy's float32 sum
        if pooled is not None:
            d = pooled[0]
            d.copy_(torch.as_tensor(det_np))
        else:
            d = torch.as_tensor(det_np).to(dev)
    if stacked:
        if pooled is not None and not marks.is_cuda:
            mk = pooled[1]
            mk.copy_(marks, non_blocking=True)
        else:
            mk = marks.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    else:
        if len(marks) != 3:
            raise ValueError("expected three (H,W,32) mark maps")
        ts = [m if isinstance(m, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(m, dtype=np.float32)) for m in marks]
        if any(t.dim() != 3 or t.shape[-1] != 32 or t.shape != ts[0].shape for t in ts):
            raise ValueError("expected three (H,W,32) mark maps")
        h, w, k = ts[0].shape
        nb = ts[0].numel() * ts[0].element_size()
        on_dev = all(t.is_cuda and t.device == dev and t.dtype == torch.float32 and t.is_contiguous() for t in ts)
        if on_dev and all(t.untyped_storage().data_ptr() == ts[0].untyped_storage().data_ptr() for t in ts) and \
                ts[1].data_ptr() == ts[0].data_ptr() + nb and ts[2].data_ptr() == ts[0].data_ptr() + 2 * nb:
            mk = torch.as_strided(ts[0], (3, h, w, k), (h * w * k, w * k, k, 1))  # consecutive slices of one device tensor: no copy
        else:
            mk = pooled[1] if pooled is not None else torch.empty((3, h, w, k), dtype=torch.float32, device=dev)
            for i in range(3):  # one H2D (or D2D) copy per map straight into its slice: no staging copy
                mk[i].copy_(ts[i], non_blocking=True)
    out = DeviceMaps(d, mk, det_sum)
    if pooled is not None:
        weakref.finalize(out, _recycle, (dev.index, int(d.shape[0]), int(d.shape[1])), pooled)
    if not reuse:
        return out
    _MAPS_CACHE[key] = out
    for a in sources:  # the entry dies with ANY of its source arrays (an address can be handed out again)
        try:
            weakref.finalize(a, _evict, key)
        except TypeError:
            pass
    if len(_MAPS_CACHE) > 2:  # bound the cache (a 2048^2 map set is 1.6 GB): drop the oldest entries
        for k in list(_MAPS_CACHE.keys())[:-2]:
            _MAPS_CACHE.pop(k, None)
    return out


# ---------------------------------------------------------------------------------------------- model translation
class TermLayout:
    """Which device column carries which named term, in the reference's key order (unit names then pair names,
    energy_graph.py:37-39)."""

    def __init__(self, spec: ModelSpec, columns: List[Tuple[str, int]], det=None, marks=None, unit=(), pair=()):
        self.spec = spec
        self.columns = columns
        self.names = [c[0] for c in columns]
        self.det, self.marks = det, marks
        self.unit, self.pair = list(unit), list(pair)
        self.max_dist = max([float(p.max_dist) for p in pair], default=0.0)

    def names_by_column(self) -> List[Optional[str]]:
        """Constructor name carried by each device column (term order of the setup); None where a term is absent."""
        out: List[Optional[str]] = [None] * len(self.spec.names)
        for name, col in self.columns:
            out[col] = name
        return out


def _check_unique(unit, pair):
    for lst in (unit, pair):  # energy_point_set.py:25-29
        for a in lst:
            for b in lst:
                if a is not b:
                    assert a.name != b.name, "does not support same name energies"
    names = [c.name for c in unit] + [c.name for c in pair]
    assert len(names) == len(set(names)), f"duplicate energy names in {names}"  # energy_graph.py:40-44


def build_layout(unit: Sequence[UnitEnergyConstructor], pair: Sequence[PairEnergyConstructor]) -> TermLayout:
    """Translates reference-style constructor lists into the device energy model."""
    unit, pair = list(unit), list(pair)
    _check_unique(unit, pair)
    for c in pair:
        if float(c.max_dist) > MAX_DEVICE_RADIUS:
            raise NotImplementedError(f"pair energy {c.name}: max_dist {c.max_dist} > {MAX_DEVICE_RADIUS} px is not supported by "
                                      f"the device cell grid")

    def only(kind, lst):
        got = [c for c in lst if type(c) is kind]
        if len(got) > 1:
            raise NotImplementedError(f"more than one {kind.__name__}")
        return got[0] if got else None

    known = (PositionEnergy, ShapeEnergy, SingleMarkEnergy, AreaPriorEnergy, RatioPriorEnergy, ConstantUnitEnergy,
             RectangleOverlapEnergy, ShapeAlignmentEnergy, DistanceIndicatorPairEnergy)
    for c in unit + pair:
        if type(c) not in known:
            raise NotImplementedError(
                f"energy term {type(c).__name__} ({c.name!r}) is a Python plug-in: only the built-in terms run on the device "
                f"(PositionEnergy, ShapeEnergy, SingleMarkEnergy, AreaPriorEnergy, RatioPriorEnergy, RectangleOverlapEnergy, "
                f"ShapeAlignmentEnergy and the toy terms ConstantUnitEnergy / DistanceIndicatorPairEnergy); there is no CPU fallback")

    toy_u, toy_p = only(ConstantUnitEnergy, unit), only(DistanceIndicatorPairEnergy, pair)
    if toy_u is not None or toy_p is not None or (not unit and not pair):
        if any(type(c) not in (ConstantUnitEnergy, DistanceIndicatorPairEnergy) for c in unit + pair):
            raise NotImplementedError("toy terms cannot be mixed with the map-driven terms")
        spec = ModelSpec(setup="toy", combinator="raw", overlap_max_dist=float(toy_p.max_dist) if toy_p else -1.0, align_max_dist=-1.0,
                         toy_unit_value=float(toy_u.value) if toy_u else 0.0, toy_pair_value=float(toy_p.value) if toy_p else 0.0,
                         toy_pair_dist=float(toy_p.max_dist) if toy_p else -1.0, toy_pair_strict=bool(toy_p.strict) if toy_p else False)
        cols = ([(toy_u.name, 0)] if toy_u else []) + ([(toy_p.name, 1)] if toy_p else [])
        return TermLayout(spec, cols, unit=unit, pair=pair)

    pos = only(PositionEnergy, unit)
    if pos is None:
        raise NotImplementedError("the device energy model needs a PositionEnergy term")
    shape = only(ShapeEnergy, unit)
    singles = [c for c in unit if type(c) is SingleMarkEnergy]
    area, ratio = only(AreaPriorEnergy, unit), only(RatioPriorEnergy, unit)
    ov, al = only(RectangleOverlapEnergy, pair), only(ShapeAlignmentEnergy, pair)
    spec = ModelSpec(pos_threshold=float(pos.threshold), combinator="raw",
                     min_area=float(area.min_area) if area else -1e300, max_area=float(area.max_area) if area else 1e300,
                     overlap_max_dist=float(ov.max_dist) if ov else -1.0, align_max_dist=float(al.max_dist) if al else -1.0,
                     rewarding=bool(al.rewarding) if al else True)
    col: Dict[int, int] = {}
    if shape is not None and not singles:
        spec.setup = "legacy"
        if ratio is not None:
            raise NotImplementedError("RatioPriorEnergy belongs to the no-calibration setup")
        if shape.param_dist_maps is not None:
            marks = shape.param_dist_maps
            spec.remap_coefs, spec.remap_intercepts = list(shape.remap_coefs), list(shape.remap_intercepts)
        elif shape.parameter_energy_map is not None:
            marks = shape.parameter_energy_map
            spec.marks_are_energies = True
        else:
            raise ValueError("ShapeEnergy needs param_dist_maps (+ remap) or parameter_energy_map")
        col = {id(pos): 0, id(shape): 1, id(ov): 2, id(al): 3, id(area): 4}
    elif len(singles) == 3 and shape is None:
        spec.setup = "nocalib"
        by = {s.param_name: s for s in singles}
        if set(by) != {"size", "ratio", "angle"}:
            raise NotImplementedError("SingleMarkEnergy terms must cover size, ratio and angle")
        order = [by["size"], by["ratio"], by["angle"]]
        if all(s.param_dist_map is not None for s in order):
            marks = [s.param_dist_map for s in order]
        elif all(s.parameter_energy_map is not None for s in order):
            marks = [s.parameter_energy_map for s in order]
            spec.marks_are_energies = True
        else:
            raise ValueError("SingleMarkEnergy needs param_dist_map or parameter_energy_map")
        spec.ratio_prior = ratio is not None
        if ratio is not None:
            spec.target_ratio = float(ratio.target_ratio)
        col = {id(pos): 0, id(order[0]): 1, id(order[1]): 2, id(order[2]): 3, id(ov): 4, id(al): 5, id(area): 6, id(ratio): 7}
    else:
        raise NotImplementedError("unsupported term set: use PositionEnergy + ShapeEnergy (legacy setup) or PositionEnergy + three "
                                  "SingleMarkEnergy (no-calibration setup), with the prior terms")
    cols = [(c.name, col[id(c)]) for c in unit + pair]
    return TermLayout(spec, cols, det=pos.detection_map, marks=marks, unit=unit, pair=pair)


def apply_combinator(layout: TermLayout, combinator) -> ModelSpec:
    """Returns the layout's spec with the given built-in combinator fused in (None -> raw sum)."""
    spec = layout.spec
    if combinator is None:
        spec.combinator, spec.comb_w, spec.comb_bias, spec.comb_threshold = "raw", [0.0] * 8, 0.0, 0.0
        return spec
    if not hasattr(combinator, "device_params"):
        raise NotImplementedError(f"{type(combinator).__name__} is a Python plug-in combinator: it is evaluated through its own "
                                  f"compute() on device-computed energy vectors, not fused into the kernels")
    if spec.setup == "toy":
        raise NotImplementedError("the toy terms take no combinator")
    kind, w, b, t = combinator.device_params(layout.names_by_column())
    spec.combinator, spec.comb_w, spec.comb_bias, spec.comb_threshold = kind, list(w), b, t
    return spec


# ---------------------------------------------------------------------------------------------- host mirror + engine
class DeviceState:
    """One device context plus the identity mirror of the Python objects stored in it."""

    def __init__(self, support_shape: Tuple[int, int], layout: TermLayout, precision: str = "fp32", device=None, reuse_maps: bool = True,
                 maps: Optional[DeviceMaps] = None):
        self.support_shape = (int(support_shape[0]), int(support_shape[1]))
        self.layout = layout
        self.precision = precision
        self.engine = Engine(self.support_shape, device=device, precision=precision)
        self.maps: Optional[DeviceMaps] = None
        if layout.det is not None:
            self.maps = maps if maps is not None else device_maps(layout.det, layout.marks, self.engine.device, reuse=reuse_maps)
            assert tuple(self.maps.det.shape) == self.support_shape, "detection map shape != support shape"
            self.engine.set_maps(self.maps.det, self.maps.marks, det_sum=self.maps.det_sum)
        self._comb_key = None
        self.engine.set_model(layout.spec)
        self.handle_of: Dict[Point, int] = {}
        self.obj_of: Dict[int, Point] = {}
        self.uid_of: Dict[Point, int] = {}
        self.by_uid: Dict[int, Point] = {}
        self._next_uid = 0
        self._kernel_key = None

    # -- model
    def use_combinator(self, combinator):
        # keyed on the VALUES sent to the device: a combinator whose weights are edited in place (training / tuning loops) is
        # sent again
        spec = apply_combinator(self.layout, combinator)
        key = None if combinator is None else (type(combinator).__name__, spec.combinator, tuple(float(v) for v in spec.comb_w),
                                                float(spec.comb_bias), float(spec.comb_threshold))
        if key != self._comb_key:
            self.engine.set_model(spec)
            self._comb_key = key

    def rebind_layout(self, layout: TermLayout):
        """Switches the energy model (EnergyGraph attached to an existing PointsSet): stored objects are re-inserted so
        that their cached unit energies follow the new model."""
        objs = self.objects()
        self.layout = layout
        if layout.det is not None:
            self.maps = device_maps(layout.det, layout.marks, self.engine.device)
            self.engine.set_maps(self.maps.det, self.maps.marks, det_sum=self.maps.det_sum)
        self._comb_key = None
        self.engine.set_model(layout.spec)
        self.engine.clear()
        self.handle_of.clear()
        self.obj_of.clear()
        self.by_uid.clear()
        self.add_many(objs)

    def use_kernels(self, intensity: float, p_kernel=None, translation_sigma: float = 2.0, max_delta: int = 8,
                    transform_sigma: float = 0.1):
        key = (float(intensity), None if p_kernel is None else tuple(float(v) for v in p_kernel), translation_sigma, max_delta,
               transform_sigma)
        if key != self._kernel_key:
            self.engine.set_kernels(intensity=intensity, p_kernel=p_kernel, translation_sigma=translation_sigma, max_delta=max_delta,
                                    transform_sigma=transform_sigma)
            self._kernel_key = key

    # -- objects
    @staticmethod
    def _marks(u: Point):
        if isinstance(u, Rectangle):
            return (float(u.size), float(u.ratio), float(u.angle))
        return (0.0, 0.0, 0.0)

    def _check_bounds(self, u: Point):
        h, w = self.support_shape
        assert 0 <= int(u.x) < h and 0 <= int(u.y) < w, "Point out of bounds"  # point_set.py:99

    def add_many(self, objs: Sequence[Point]):
        objs = list(objs)
        if not objs:
            return
        for u in objs:
            self._check_bounds(u)
        xy = np.array([[int(u.x), int(u.y)] for u in objs], dtype=np.int32)
        marks = np.array([self._marks(u) for u in objs], dtype=np.float64)
        uids = []
        for u in objs:
            if u not in self.uid_of:
                self.uid_of[u] = self._next_uid
                self._next_uid += 1
            uids.append(self.uid_of[u])
        classes = classes_of_marks(marks)
        handles = self.engine.add_objects(xy, marks, classes=classes, uid=np.array(uids, dtype=np.int64))
        for u, h in zip(objs, handles):
            self.handle_of[u] = int(h)
            self.obj_of[int(h)] = u
            self.by_uid[self.uid_of[u]] = u

    def add(self, u: Point):
        if u in self.handle_of:
            return  # a set: adding a member again is a no-op (point_set.py:102-103)
        self.add_many([u])

    def remove(self, u: Point):
        h = self.handle_of.pop(u)  # KeyError for an unknown object, like set.remove (point_set.py:105-106)
        self.obj_of.pop(h, None)
        self.by_uid.pop(self.uid_of.get(u, -1), None)
        self.engine.remove_objects([h])

    def __len__(self):
        return len(self.handle_of)

    def __contains__(self, u):
        return u in self.handle_of

    def objects(self) -> List[Point]:
        """Cell-major order (PointsSetIterator, point_set.py:12-42); inside a cell: slot order."""
        return [self.obj_of[h] for h in sorted(self.obj_of)]

    def handles(self, objs: Sequence[Point]) -> np.ndarray:
        return np.array([self.handle_of[u] for u in objs], dtype=np.uint32)

    def copy(self) -> "DeviceState":
        new = DeviceState.__new__(DeviceState)
        new.support_shape, new.layout, new.precision = self.support_shape, self.layout, self.precision
        new.engine = Engine(self.support_shape, device=self.engine.device, precision=self.precision)
        new.maps = self.maps
        if self.maps is not None:
            new.engine.set_maps(self.maps.det, self.maps.marks, det_sum=self.maps.det_sum)
        new._comb_key = None
        new.engine.set_model(self.layout.spec)
        new.engine.copy_state_from(self.engine)
        new.handle_of = dict(self.handle_of)
        new.obj_of = dict(self.obj_of)
        new.uid_of = dict(self.uid_of)
        new.by_uid = dict(self.by_uid)
        new._next_uid = self._next_uid
        new._kernel_key = None
        return new

    def refresh_from_device(self):
        """Rebuilds the host mirror after a device-side chain changed the configuration.  Objects that survived keep
        their Python identity (matched by uid); new ones become fresh Rectangle / Point instances."""
        handles, xy, marks, uid = self.engine.read_objects()
        by_uid = {v: k for k, v in self.uid_of.items()}
        self.handle_of.clear()
        self.obj_of.clear()
        new_uid_of: Dict[Point, int] = {}
        is_toy = self.layout.spec.setup == "toy"
        for h, p, m, i in zip(handles, xy, marks, uid):
            old = by_uid.get(int(i))
            same = old is not None and int(old.x) == int(p[0]) and int(old.y) == int(p[1]) and (
                is_toy or not isinstance(old, Rectangle) or
                (np.float32(old.size) == np.float32(m[0]) and np.float32(old.ratio) == np.float32(m[1]) and np.float32(old.angle) == np.float32(m[2])))
            if same:
                u = old
            else:
                u = Point(int(p[0]), int(p[1])) if is_toy else Rectangle(int(p[0]), int(p[1]), float(m[0]), float(m[1]), float(m[2]))
            self.handle_of[u] = int(h)
            self.obj_of[int(h)] = u
            new_uid_of[u] = int(i)
        self.uid_of = new_uid_of
        self.by_uid = {i: u for u, i in new_uid_of.items()}
        self._next_uid = max([self._next_uid] + [int(i) + 1 for i in uid])

    # -- perturbations
    def proposal_record(self, removal: Optional[Point], addition: Optional[Point], kernel: int = 0, delta=(0.0, 0.0), param_id: int = 0,
                        new_class: int = 0, u: float = 0.5) -> np.ndarray:
        p = np.zeros(1, dtype=_lib.PROPOSAL_DTYPE)
        p["kernel"] = kernel
        if removal is not None:
            if removal not in self.handle_of:
                raise KeyError(removal)  # energy_point_set.py:88-100
            p["rem_x"], p["rem_y"], p["rem_uid"] = int(removal.x), int(removal.y), self.uid_of[removal]
        else:
            p["rem_uid"] = _lib.NO_OBJECT
        if addition is not None:
            self._check_bounds(addition)
            m = np.array([self._marks(addition)], dtype=np.float64)
            c = classes_of_marks(m)[0]
            p["add_x"], p["add_y"] = int(addition.x), int(addition.y)
            p["add_size"], p["add_ratio"], p["add_angle"] = m[0]
            p["add_cls"] = int(c[0]) | (int(c[1]) << 8) | (int(c[2]) << 16)
            if addition not in self.uid_of:
                self.uid_of[addition] = self._next_uid
                self._next_uid += 1
            p["add_uid"] = self.uid_of[addition]
        else:
            p["add_uid"] = _lib.NO_OBJECT
        p["delta0"], p["delta1"] = float(delta[0]), float(delta[1])
        p["param_id"], p["new_class"], p["u"] = int(param_id), int(new_class), float(u)
        return p
