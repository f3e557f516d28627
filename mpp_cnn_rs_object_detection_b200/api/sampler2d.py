"""utils/sampler2d.py: `sample_point_2d`, same signature; the weighted draws run on the device (mpp_sample_points_2d: row prefix
sums + inverse CDF, replacing `rng.choice` over all H*W pixels, O(H*W) per call in the reference)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from .. import _lib


def _device_draws(density, img_shape, n: int, seed: int) -> np.ndarray:
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("sample_point_2d needs a CUDA device (no CPU fallback)")
    dev = density.device if isinstance(density, torch.Tensor) and density.is_cuda else torch.device("cuda", torch.cuda.current_device())
    h, w = int(img_shape[0]), int(img_shape[1])
    d = density if isinstance(density, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(density, dtype=np.float32))
    d = d.to(device=dev, dtype=torch.float32).contiguous()
    assert tuple(d.shape) == (h, w), "density shape != img_shape"
    out = torch.empty((n, 2), dtype=torch.int32, device=dev)
    scratch = torch.empty(h * (w + 1) + h, dtype=torch.float64, device=dev)
    _lib.check(lib.mpp_sample_points_2d(d.data_ptr(), h, w, n, int(seed) & 0xFFFFFFFFFFFFFFFF, out.data_ptr(), scratch.data_ptr(), dev.index,
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out.cpu().numpy().astype(np.int64)


def sample_point_2d(img_shape: Tuple[int, int], size: int = 1, density: np.ndarray = None, skip_normalization: bool = False,
                    rng: np.random.Generator = None, mask: np.ndarray = None) -> np.ndarray:
    """Samples `size` point(s) of a rectangular image according to a density (uniformly when None): coordinates (size, 2).
    As in the reference (sampler2d.py:24-46): without a density and without a mask the two coordinates are drawn independently
    on the host; with a mask only, the mask is the weight map; with a density, `density[mask] = 0` excludes the masked pixels,
    and the draws are WITHOUT replacement.  The random stream is Philox keyed by a seed taken from `rng` (not numpy's)."""
    if rng is None:
        rng = np.random.default_rng()
    if density is None and mask is None:
        return np.array([rng.choice(np.arange(0, img_shape[0]), size=size), rng.choice(np.arange(0, img_shape[1]), size=size)]).T
    if density is None:
        weights = np.asarray(mask, dtype=np.float32)
    else:
        weights = density.detach().float().cpu().numpy().copy() if isinstance(density, torch.Tensor) else np.array(density, dtype=np.float32)
        if mask is not None:
            weights[np.asarray(mask)] = 0
    if not skip_normalization or mask is not None:
        total = float(np.sum(weights, dtype=np.float64))
        if not total > 0:
            raise ValueError("probabilities do not sum to a positive value")
    if int(np.count_nonzero(weights)) < size:
        raise ValueError("Cannot take a larger sample than population when replace is False")
    # sequential sampling without replacement == keep the first occurrence of every pixel and redraw the rest
    chosen, seen = [], set()
    while len(chosen) < size:
        draws = _device_draws(weights, img_shape, size - len(chosen), int(rng.integers(0, 2 ** 62)))
        for x, y in draws:
            if (int(x), int(y)) not in seen:
                seen.add((int(x), int(y)))
                chosen.append((int(x), int(y)))
    return np.array(chosen, dtype=np.int64).reshape(size, 2)
