"""Synthetic scenes for the benchmark configs (BASELINE.json configs[2:5]).

Objects follow the mark distributions of the reference's generator (data/make_synth_data.py:16-31:
size ~ N(8,1), ratio ~ clip(N(.5,.1),.1,1), angle ~ U(0,pi), uniform integer centres, candidates that
intersect an already kept rectangle are rejected).  The sampler's inputs are maps, not images, so instead
of rendering an image and running the UNets the maps are synthesised directly:

  detection map  (H,W)    f32 : 0.02 background + a Gaussian blob (sigma 2 px, peak ~1) per object
  mark maps   3x(H,W,32)  f32 : rows sum to 1; peaked (0.8) at the true class within +-4 px of a centre,
                                random elsewhere

`make_scene` builds everything with numpy (tests, golden vectors); `make_scene_torch` builds the maps on a
torch device for the large benchmark scenes (2048^2 marks are 1.5 GB).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

N_CLASSES = 32
MARK_RANGES = ((0.0, 32.0), (0.0, 1.0), (0.0, math.pi))  # models/shape_net/shape_net_model.py:80-85


def _corners(x, y, size, ratio, angle):
    length = (2 * size) / (1 + ratio)
    width = ratio * length
    a = angle + math.pi / 2
    c, s = math.cos(a), math.sin(a)
    pts = []
    for lx, ly in ((length / 2, width / 2), (length / 2, -width / 2), (-length / 2, -width / 2), (-length / 2, width / 2)):
        pts.append((lx * c - ly * s + x, lx * s + ly * c + y))
    return pts


def _sat_disjoint(p, q) -> bool:
    """Separating-axis test for two convex quads; True when their interiors do not intersect."""
    for poly in (p, q):
        for i in range(4):
            ex, ey = poly[(i + 1) % 4][0] - poly[i][0], poly[(i + 1) % 4][1] - poly[i][1]
            nx, ny = -ey, ex
            a = [nx * v[0] + ny * v[1] for v in p]
            b = [nx * v[0] + ny * v[1] for v in q]
            if max(a) <= min(b) or max(b) <= min(a):
                return True
    return False


def make_objects(seed: int, shape: Tuple[int, int], n_rect: int) -> np.ndarray:
    """Returns (N,5) float64 rows (x, y, size, ratio, angle) of mutually non-intersecting rectangles."""
    rng = np.random.default_rng(seed)
    h, w = shape
    xs = rng.integers(0, h, n_rect)
    ys = rng.integers(0, w, n_rect)
    sizes = rng.normal(8, 1.0, n_rect)
    ratios = np.clip(rng.normal(0.5, 0.1, n_rect), 0.1, 1)
    angles = rng.uniform(0, np.pi, n_rect)
    cell = 32
    grid = {}
    kept: List[Tuple] = []
    for k in range(n_rect):
        x, y = int(xs[k]), int(ys[k])
        poly = _corners(x, y, sizes[k], ratios[k], angles[k])
        ci, cj = x // cell, y // cell
        ok = True
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                for other in grid.get((ci + di, cj + dj), ()):
                    if not _sat_disjoint(poly, other):
                        ok = False
                        break
                if not ok:
                    break
            if not ok:
                break
        if ok:
            grid.setdefault((ci, cj), []).append(poly)
            kept.append((x, y, sizes[k], ratios[k], angles[k]))
    return np.array(kept, dtype=np.float64).reshape(-1, 5)


def mark_classes(objs: np.ndarray) -> np.ndarray:
    """(N,3) int classes of (size, ratio, angle): max{c : v >= edge_c} (mappings.py:45-61) in float64."""
    out = np.zeros((len(objs), 3), dtype=np.int64)
    for i, (lo, hi) in enumerate(MARK_RANGES):
        edges = np.linspace(lo, hi, N_CLASSES + 1)[:-1]
        out[:, i] = np.clip(np.searchsorted(edges, objs[:, 2 + i], side="right") - 1, 0, N_CLASSES - 1)
    return out


def make_maps(objs: np.ndarray, shape: Tuple[int, int], seed: int, blob_sigma: float = 2.0, peak: float = 0.8, peak_radius: int = 4):
    """Detection map and mark maps consistent with the given objects ((N,5) rows x, y, size, ratio, angle); the random
    background is drawn from default_rng(seed + 1).  Returns (det (H,W) f32, marks [3 x (H,W,32) f32])."""
    h, w = shape
    rng = np.random.default_rng(seed + 1)
    det = np.full((h, w), 0.02, dtype=np.float32)
    r = int(math.ceil(3 * blob_sigma))
    ax = np.arange(-r, r + 1)
    blob = np.exp(-(ax[:, None] ** 2 + ax[None, :] ** 2) / (2 * blob_sigma ** 2)).astype(np.float32)
    for o in objs:
        x, y = int(o[0]), int(o[1])
        x0, x1, y0, y1 = max(0, x - r), min(h, x + r + 1), max(0, y - r), min(w, y + r + 1)
        det[x0:x1, y0:y1] += 0.97 * blob[x0 - x + r:x1 - x + r, y0 - y + r:y1 - y + r]
    det = np.clip(det, 0.0, 0.999).astype(np.float32)
    cls = mark_classes(objs)
    marks = []
    for i in range(3):
        m = rng.random((h, w, N_CLASSES), dtype=np.float32)
        m = m * m + np.float32(1e-3)
        m /= m.sum(-1, keepdims=True)
        for o, c in zip(objs, cls):
            x, y = int(o[0]), int(o[1])
            x0, x1, y0, y1 = max(0, x - peak_radius), min(h, x + peak_radius + 1), max(0, y - peak_radius), min(w, y + peak_radius + 1)
            win = m[x0:x1, y0:y1]
            rest = 1.0 - win[..., c[i]]
            win *= ((1.0 - peak) / np.maximum(rest, 1e-6))[..., None]
            win[..., c[i]] = peak
            win /= win.sum(-1, keepdims=True)
        marks.append(np.ascontiguousarray(m, dtype=np.float32))
    return det, marks


def make_scene(seed: int, shape: Tuple[int, int], n_rect: int, blob_sigma: float = 2.0, peak: float = 0.8,
               peak_radius: int = 4):
    """numpy scene: returns (objects (N,5) f64, det (H,W) f32, marks [3 x (H,W,32) f32])."""
    objs = make_objects(seed, shape, n_rect)
    det, marks = make_maps(objs, shape, seed, blob_sigma, peak, peak_radius)
    return objs, det, marks


def make_scene_torch(seed: int, shape: Tuple[int, int], n_rect: int, device, blob_sigma: float = 2.0,
                     peak: float = 0.8, peak_radius: int = 4):
    """Same recipe with the maps built on `device` (torch); returns (objects (N,5) f64 numpy, det, marks (3,H,W,32))."""
    import torch

    objs = make_objects(seed, shape, n_rect)
    h, w = shape
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    det = torch.full((h, w), 0.02, dtype=torch.float32, device=device)
    r = int(math.ceil(3 * blob_sigma))
    n = len(objs)
    cls = torch.as_tensor(mark_classes(objs), device=device)
    if n > 0:
        cx = torch.as_tensor(objs[:, 0].astype(np.int64), device=device)
        cy = torch.as_tensor(objs[:, 1].astype(np.int64), device=device)
        ax = torch.arange(-r, r + 1, device=device)
        dx, dy = torch.meshgrid(ax, ax, indexing="ij")
        blob = (0.97 * torch.exp(-(dx ** 2 + dy ** 2).float() / (2 * blob_sigma ** 2))).reshape(1, -1)
        px = (cx[:, None] + dx.reshape(1, -1))
        py = (cy[:, None] + dy.reshape(1, -1))
        ok = (px >= 0) & (px < h) & (py >= 0) & (py < w)
        det.view(-1).index_put_(((px * w + py)[ok],), blob.expand(n, -1)[ok], accumulate=True)
    det.clamp_(0.0, 0.999)
    marks = torch.empty((3, h, w, N_CLASSES), dtype=torch.float32, device=device)
    for i in range(3):
        m = marks[i]
        rows = max(1, (1 << 26) // (w * N_CLASSES))
        for s in range(0, h, rows):
            blk = torch.rand((min(rows, h - s), w, N_CLASSES), generator=gen, device=device, dtype=torch.float32)
            blk = blk * blk + 1e-3
            blk /= blk.sum(-1, keepdim=True)
            m[s:s + blk.shape[0]] = blk
        if n > 0:
            ax = torch.arange(-peak_radius, peak_radius + 1, device=device)
            dx, dy = torch.meshgrid(ax, ax, indexing="ij")
            px = (cx[:, None] + dx.reshape(1, -1))
            py = (cy[:, None] + dy.reshape(1, -1))
            ok = (px >= 0) & (px < h) & (py >= 0) & (py < w)
            pix = (px * w + py)[ok]
            c = cls[:, i][:, None].expand(-1, px.shape[1])[ok]
            flat = m.view(-1, N_CLASSES)
            win = flat[pix]
            rest = 1.0 - win.gather(1, c[:, None]).squeeze(1)
            win *= ((1.0 - peak) / rest.clamp_min(1e-6))[:, None]
            win.scatter_(1, c[:, None], peak)
            win /= win.sum(-1, keepdim=True)
            flat[pix] = win
    return objs, det, marks


BAND_BLOCK = 64  # rows per independently seeded block of the mark-map background (band generator)


def det_map_numpy(objs: np.ndarray, shape: Tuple[int, int], blob_sigma: float = 2.0) -> np.ndarray:
    """Detection map of a scene with a FIXED accumulation order (np.add.at), so that every rank of a split scene holds bit-equal
    copies of the rows it shares with its neighbours (the atomics of the torch builder may round overlapping blobs differently)."""
    h, w = shape
    det = np.full(h * w, 0.02, dtype=np.float32)
    r = int(math.ceil(3 * blob_sigma))
    ax = np.arange(-r, r + 1)
    dx, dy = np.meshgrid(ax, ax, indexing="ij")
    blob = (0.97 * np.exp(-(dx ** 2 + dy ** 2).astype(np.float32) / np.float32(2 * blob_sigma ** 2))).astype(np.float32).reshape(1, -1)
    if len(objs):
        px = objs[:, 0].astype(np.int64)[:, None] + dx.reshape(1, -1)
        py = objs[:, 1].astype(np.int64)[:, None] + dy.reshape(1, -1)
        ok = (px >= 0) & (px < h) & (py >= 0) & (py < w)
        np.add.at(det, (px * w + py)[ok], np.broadcast_to(blob, px.shape)[ok])
    return np.clip(det, 0.0, 0.999).astype(np.float32).reshape(h, w)


def make_scene_band_torch(seed: int, shape: Tuple[int, int], n_rect: int, device, row0: int = 0, rows: int = None,
                          peak: float = 0.8, peak_radius: int = 4, objs: np.ndarray = None, det: np.ndarray = None):
    """The rows [row0, row0 + rows) of a scene whose maps are a function of (seed, pixel) only, whatever the band: the mark-map
    background comes from one generator per 64-row block and the peaked windows are assigned pixel by pixel to the
    lowest-numbered object within reach (order-independent), so ranks that hold overlapping bands of one scene hold
    bit-equal rows.  Returns (objects of the WHOLE scene (N,5) f64, det of the whole scene (H,W) f32 numpy,
    marks of the band (3, rows, W, 32) on `device`)."""
    import torch

    h, w = shape
    rows = h - row0 if rows is None else rows
    if objs is None:
        objs = make_objects(seed, shape, n_rect)
    if det is None:
        det = det_map_numpy(objs, shape)
    n = len(objs)
    cls = torch.as_tensor(mark_classes(objs), device=device)
    marks = torch.empty((3, rows, w, N_CLASSES), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    # lowest-numbered object whose (2 r + 1)^2 window covers each pixel of the band
    owner = torch.full((rows * w,), n, dtype=torch.int64, device=device)
    if n > 0:
        near = np.nonzero((objs[:, 0] >= row0 - peak_radius) & (objs[:, 0] < row0 + rows + peak_radius))[0]
        if len(near):
            idx = torch.as_tensor(near, device=device)
            cx = torch.as_tensor(objs[near, 0].astype(np.int64), device=device)
            cy = torch.as_tensor(objs[near, 1].astype(np.int64), device=device)
            ax = torch.arange(-peak_radius, peak_radius + 1, device=device)
            dx, dy = torch.meshgrid(ax, ax, indexing="ij")
            px = cx[:, None] + dx.reshape(1, -1) - row0
            py = cy[:, None] + dy.reshape(1, -1)
            ok = (px >= 0) & (px < rows) & (py >= 0) & (py < w)
            owner.scatter_reduce_(0, (px * w + py)[ok], idx[:, None].expand(-1, px.shape[1])[ok], reduce="amin")
    has = owner < n
    pix = torch.nonzero(has).squeeze(1)
    for i in range(3):
        m = marks[i]
        b0, b1 = row0 // BAND_BLOCK, (row0 + rows + BAND_BLOCK - 1) // BAND_BLOCK
        for b in range(b0, b1):
            gen.manual_seed((seed * 1000003 + i * 7919 + b * 104729 + 12345) & 0x7FFFFFFFFFFF)
            blk = torch.rand((BAND_BLOCK, w, N_CLASSES), generator=gen, device=device, dtype=torch.float32)
            blk = blk * blk + 1e-3
            blk /= blk.sum(-1, keepdim=True)
            s0, s1 = max(b * BAND_BLOCK, row0), min((b + 1) * BAND_BLOCK, row0 + rows, h)
            m[s0 - row0:s1 - row0] = blk[s0 - b * BAND_BLOCK:s1 - b * BAND_BLOCK]
        if len(pix):
            c = cls[owner[pix], i]
            flat = m.view(-1, N_CLASSES)
            win = flat[pix]
            rest = 1.0 - win.gather(1, c[:, None]).squeeze(1)
            win *= ((1.0 - peak) / rest.clamp_min(1e-6))[:, None]
            win.scatter_(1, c[:, None], peak)
            win /= win.sum(-1, keepdim=True)
            flat[pix] = win
    return objs, det, marks


def render_tiles_torch(objs_per_tile, shape: Tuple[int, int], device, seed: int = 0):
    """Synthetic RGB tiles (B, 3, H, W) in [0, 1]: noise background + a bright blob per object (what data/make_synth_data.py
    renders as filled rectangles); the input of the map-producing CNNs in the tile benchmark."""
    import torch

    h, w = shape
    b = len(objs_per_tile)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 77)
    img = 0.2 * torch.rand((b, 3, h, w), generator=gen, device=device, dtype=torch.float32)
    for k, objs in enumerate(objs_per_tile):
        if len(objs):
            det = torch.as_tensor(det_map_numpy(objs, shape, blob_sigma=3.0), device=device)
            img[k] += 0.8 * det.unsqueeze(0)
    return img.clamp_(0.0, 1.0)


def inject_objects_torch(det, marks, objs_per_tile, peak: float = 0.8, peak_radius: int = 4, blob_sigma: float = 2.0):
    """Writes a synthetic scene into CNN-produced maps, on the device and in place: randomly initialised networks detect nothing,
    so the benchmark keeps their outputs as the background (scaled to the 0.02 level of the synthetic recipe) and adds, per
    object, the detection blob and the peaked mark rows of `make_maps`.  det (B, H, W), marks (B, 3, H, W, 32)."""
    import torch

    bsz, h, w = det.shape
    device = det.device
    det.mul_(0.04)
    r = int(math.ceil(3 * blob_sigma))
    ax = torch.arange(-r, r + 1, device=device)
    dx, dy = torch.meshgrid(ax, ax, indexing="ij")
    blob = (0.97 * torch.exp(-(dx ** 2 + dy ** 2).float() / (2 * blob_sigma ** 2))).reshape(1, -1)
    axp = torch.arange(-peak_radius, peak_radius + 1, device=device)
    dxp, dyp = torch.meshgrid(axp, axp, indexing="ij")
    for k, objs in enumerate(objs_per_tile):
        n = len(objs)
        if n == 0:
            continue
        cx = torch.as_tensor(objs[:, 0].astype(np.int64), device=device)
        cy = torch.as_tensor(objs[:, 1].astype(np.int64), device=device)
        px, py = cx[:, None] + dx.reshape(1, -1), cy[:, None] + dy.reshape(1, -1)
        ok = (px >= 0) & (px < h) & (py >= 0) & (py < w)
        det[k].view(-1).index_put_(((px * w + py)[ok],), blob.expand(n, -1)[ok], accumulate=True)
        cls = torch.as_tensor(mark_classes(objs), device=device)
        px, py = cx[:, None] + dxp.reshape(1, -1), cy[:, None] + dyp.reshape(1, -1)
        ok = (px >= 0) & (px < h) & (py >= 0) & (py < w)
        owner = torch.full((h * w,), n, dtype=torch.int64, device=device)
        owner.scatter_reduce_(0, (px * w + py)[ok], torch.arange(n, device=device)[:, None].expand(-1, px.shape[1])[ok], reduce="amin")
        pix = torch.nonzero(owner < n).squeeze(1)
        for i in range(3):
            flat = marks[k, i].view(-1, N_CLASSES)
            c = cls[owner[pix], i]
            win = flat[pix]
            rest = 1.0 - win.gather(1, c[:, None]).squeeze(1)
            win *= ((1.0 - peak) / rest.clamp_min(1e-6))[:, None]
            win.scatter_(1, c[:, None], peak)
            win /= win.sum(-1, keepdim=True)
            flat[pix] = win
    det.clamp_(0.0, 0.999)
    return det, marks
