"""Host-side sharding of the sampler across the GPUs of one box (one process per GPU, torch.distributed).

Decompositions:

* one scene split into row bands with PEER ACCESS (`PeerSplitScene`, the production path of BASELINE configs[3]): every rank
  runs ONE persistent dataflow kernel over the windows of its band (mpp_run_windows_batch); the boundary cells of the
  neighbour bands are read and written directly in the neighbours' device memory over NVLink (CUDA IPC mappings) and window
  completions are stamped into the neighbours' completion grids, so there is no exchange step, no host synchronisation and no
  collective on the data path.  torch.distributed only carries the 192-byte IPC handles at set-up;

* independent tiles / images (`shard_items`): each rank samples its own images; no data-path collective, results are
  gathered on the host (the reference's own decomposition: mpp_model.py:231-264 maps patches over a process pool);
* one scene split into row bands (`SplitScene`): each rank owns the window rows that start inside its band and keeps a
  copy of the objects up to 96 px beyond it.  Window rows of equal colour ci = wi mod 3 are >= 65 px apart, so all ranks
  run one colour-row phase (mpp_run_window_rows) concurrently and exchange their boundary objects with their two
  neighbours between phases: 3 exchanges of a few KB per sweep (latency-bound; NCCL send/recv over NVLink, or any other
  transport).  Grid offsets, random streams and uids depend only on (seed, sweep, window), so the split chain is the
  same chain as mpp_run_windows(schedule='colours') on a single GPU, bit for bit.

The transport is pluggable so that the protocol is testable without GPUs (gloo, mock engine) and on a single GPU
(several bands in one process)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

HALO = 64          # reach of a Delta-energy: 32 px to an affected object + 32 px to its partners
BAND_ALIGN = 32    # band boundaries are multiples of the cell size
RECORD = 8         # doubles per packed object (mpp_pack_rows)


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pins the calling process (and therefore the pinned host buffers it allocates afterwards: first touch) to the CPUs of the
    NUMA node the GPU hangs off, so that with one process per GPU the host-to-device uploads of all ranks do not cross the
    socket interconnect.  Returns the CPU list, or None when the topology is not exposed (then nothing is changed)."""
    import os
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = []
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus += list(range(int(lo), int(hi) + 1))
            elif part:
                cpus.append(int(part))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except (OSError, AttributeError, ValueError):
        return None


def shard_items(n_items: int, world: int, rank: int) -> List[int]:
    """Round-robin assignment of independent tiles / images to ranks."""
    return list(range(rank, n_items, world))


def row_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """`world` contiguous row bands [r0, r1) covering [0, height), boundaries aligned to 32 px, sizes as equal as possible."""
    n_cells = (height + BAND_ALIGN - 1) // BAND_ALIGN
    if world > n_cells:
        raise ValueError(f"cannot split {height} rows into {world} bands of at least {BAND_ALIGN} px")
    cuts = [round(i * n_cells / world) * BAND_ALIGN for i in range(world + 1)]
    cuts[-1] = height
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def grid_offset(seed: int, sweep_id: int) -> Tuple[int, int]:
    """(ox, oy) of the window grid of a sweep: the host-side twin of mpp_window_grid."""
    h = _splitmix64((seed & _M64) ^ _splitmix64(sweep_id & _M64))
    return h & 31, (h >> 5) & 31


def first_grid_line_at_or_after(row: int, ox: int) -> int:
    """Smallest 32*k - ox >= row: first pixel row of the first window row that starts at or after `row`."""
    k = -(-(row + ox) // 32)
    return 32 * k - ox


class SplitScene:
    """One rank's share of a scene split into row bands.  `engine` must expose run_window_rows / window_grid / pack_rows /
    unpack_rows / read_objects / add_objects (mpp_cnn_rs_object_detection_b200.engine.Engine does)."""

    def __init__(self, engine, height: int, rank: int, world: int, capacity: int = 8192):
        self.engine, self.height, self.rank, self.world = engine, int(height), rank, world
        self.r0, self.r1 = row_bands(height, world)[rank]
        self.capacity = capacity

    # ---- initial distribution / final collection
    def local_rows(self) -> Tuple[int, int]:
        """Rows of the objects this rank must hold: its band plus the halo it reads and the strip it may come to own."""
        return max(0, self.r0 - HALO - BAND_ALIGN), min(self.height, self.r1 + HALO + BAND_ALIGN)

    def select_initial(self, xy: np.ndarray) -> np.ndarray:
        lo, hi = self.local_rows()
        return (xy[:, 0] >= lo) & (xy[:, 0] < hi)

    def owned_objects(self):
        """(xy, marks, uid) of the objects whose row lies in this rank's band (valid after the last exchange of a sweep)."""
        _, xy, marks, uid = self.engine.read_objects()
        sel = (xy[:, 0] >= self.r0) & (xy[:, 0] < self.r1)
        return xy[sel], marks[sel], uid[sel]

    # ---- one colour-row phase
    def compute(self, ci: int, per_visit: int, n_warps: int, temperature: float, seed: int, sweep_id: int):
        self.engine.run_window_rows(per_visit, n_warps, temperature, seed, sweep_id, ci, self.r0, self.r1)

    def boundaries(self, seed: int, sweep_id: int) -> Tuple[Optional[int], Optional[int]]:
        """Pixel rows (g_up, g_down) where ownership switches to the upper / lower neighbour in this sweep."""
        ox, _ = self.engine.window_grid(seed, sweep_id)
        g_up = first_grid_line_at_or_after(self.r0, ox) if self.rank > 0 else None
        g_down = first_grid_line_at_or_after(self.r1, ox) if self.rank < self.world - 1 else None
        return g_up, g_down

    def pack(self, seed: int, sweep_id: int):
        """Messages for (upper neighbour, lower neighbour): fixed-size [capacity + 1, 8] float64 device tensors whose first row
        holds the record count.  Upper neighbour receives my rows [g_up, r0 + 96); lower receives [r1 - 64, g_down)."""
        g_up, g_down = self.boundaries(seed, sweep_id)
        up = self._pack(g_up, self.r0 + HALO + BAND_ALIGN) if g_up is not None else None
        down = self._pack(self.r1 - HALO, g_down) if g_down is not None else None
        return up, down

    def _pack(self, lo: int, hi: int) -> torch.Tensor:
        rec = self.engine.pack_rows(lo, hi, capacity=self.capacity)
        msg = torch.zeros((self.capacity + 1, RECORD), dtype=torch.float64, device=rec.device)
        msg[0, 0] = float(len(rec))
        msg[1:1 + len(rec)] = rec
        return msg

    def unpack(self, from_up: Optional[torch.Tensor], from_down: Optional[torch.Tensor], seed: int, sweep_id: int):
        """Replaces my copies of the neighbours' boundary rows: [r0 - 64, g_up) from above, [g_down, r1 + 96) from below."""
        g_up, g_down = self.boundaries(seed, sweep_id)
        if from_up is not None:
            n = int(from_up[0, 0].item())
            self.engine.unpack_rows(self.r0 - HALO, g_up, from_up[1:1 + n])
        if from_down is not None:
            n = int(from_down[0, 0].item())
            self.engine.unpack_rows(g_down, self.r1 + HALO + BAND_ALIGN, from_down[1:1 + n])


class PeerSplitScene:
    """One rank's band of a scene sampled with peer access to the neighbour bands (see the module docstring).  `engine` is an
    Engine of the WHOLE scene's shape; its maps may be band-local (Engine.set_maps_band over `map_rows()`)."""

    MIN_BAND = 384
    MAP_MARGIN = 64   # map rows a rank needs beyond its band: a window starting in the band reaches 31 rows below it, a
                      # data-driven translation 8 more

    def __init__(self, engine, height: int, rank: int, world: int):
        self.engine, self.height, self.rank, self.world = engine, int(height), int(rank), int(world)
        bands = row_bands(height, world)
        if world > 1 and min(b[1] - b[0] for b in bands) < self.MIN_BAND:
            raise ValueError(f"{height} rows split {world} ways gives bands under {self.MIN_BAND} rows")
        self.r0, self.r1 = bands[rank]
        self.attached = False

    @staticmethod
    def map_rows(height: int, rank: int, world: int) -> Tuple[int, int]:
        r0, r1 = row_bands(height, world)[rank]
        return max(0, r0 - PeerSplitScene.MAP_MARGIN), min(height, r1 + PeerSplitScene.MAP_MARGIN)

    def select_initial(self, xy: np.ndarray) -> np.ndarray:
        """Objects this rank stores: those whose 32-px cell row lies in its band (bands are whole cell rows)."""
        return (xy[:, 0] >= self.r0) & (xy[:, 0] < self.r1)

    def attach_dist(self, group=None):
        """Exchanges the CUDA IPC handles of the contexts with the two neighbour ranks and maps their state (host plumbing over
        torch.distributed; ends with a barrier: nobody may run before everybody is attached)."""
        import torch.distributed as dist
        mine = self.engine.split_export()
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
        else:
            handles = [mine]
        up = handles[self.rank - 1] if self.rank > 0 else None
        down = handles[self.rank + 1] if self.rank < self.world - 1 else None
        self.engine.split_attach(self.r0, self.r1, up, down)
        self.attached = True
        if self.world > 1:
            dist.barrier(group=group)

    def attach_local(self, up_engine, down_engine):
        self.engine.split_attach_local(self.r0, self.r1, up_engine, down_engine)
        self.attached = True

    def run(self, n_sweeps: int, per_visit: int, n_warps: int, t0: float, seed: int, alpha_t: float = 1.0, t_target: float = 0.0,
            sweep_offset: int = 0, max_ctas: int = 0, read_counters: bool = False):
        """`n_sweeps` sweeps of this rank's band: one asynchronous launch.  Every rank must issue the same calls."""
        from .engine import run_windows_batch
        assert self.attached, "attach_dist() / attach_local() first"
        return run_windows_batch([self.engine], [seed], n_sweeps, per_visit, n_warps=n_warps, t0=t0, alpha_t=alpha_t, t_target=t_target,
                                 sweep_offset=sweep_offset, max_ctas=max_ctas, read_counters=read_counters)

    def owned_objects(self):
        """(xy, marks, uid) of this rank's band."""
        _, xy, marks, uid = self.engine.read_objects()
        return xy, marks, uid

    def detach(self):
        if self.attached:
            self.engine.split_detach()
            self.attached = False


# ------------------------------------------------------------------------------------------------ transports
def exchange_dist(scene: SplitScene, up: Optional[torch.Tensor], down: Optional[torch.Tensor], group=None):
    """Neighbour exchange over torch.distributed (NCCL on GPUs, gloo in the CPU tests): returns (from_up, from_down)."""
    import torch.distributed as dist
    ops, from_up, from_down = [], None, None
    if scene.rank > 0:
        from_up = torch.empty_like(up)
        ops += [dist.P2POp(dist.isend, up, scene.rank - 1, group), dist.P2POp(dist.irecv, from_up, scene.rank - 1, group)]
    if scene.rank < scene.world - 1:
        from_down = torch.empty_like(down)
        ops += [dist.P2POp(dist.isend, down, scene.rank + 1, group), dist.P2POp(dist.irecv, from_down, scene.rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return from_up, from_down


def sweep_dist(scene: SplitScene, per_visit: int, n_warps: int, temperature: float, seed: int, sweep_id: int, group=None):
    """One sweep of a split scene on this rank (call on every rank)."""
    for ci in range(3):
        scene.compute(ci, per_visit, n_warps, temperature, seed, sweep_id)
        up, down = scene.pack(seed, sweep_id)
        from_up, from_down = exchange_dist(scene, up, down, group)
        scene.unpack(from_up, from_down, seed, sweep_id)


def sweep_local(scenes: Sequence[SplitScene], per_visit: int, n_warps: int, temperature: float, seed: int, sweep_id: int):
    """The same protocol for all bands held in ONE process (single-GPU emulation used by the tests)."""
    for ci in range(3):
        for s in scenes:
            s.compute(ci, per_visit, n_warps, temperature, seed, sweep_id)
        msgs = [s.pack(seed, sweep_id) for s in scenes]
        for r, s in enumerate(scenes):
            from_up = msgs[r - 1][1] if r > 0 else None
            from_down = msgs[r + 1][0] if r < len(scenes) - 1 else None
            s.unpack(from_up, from_down, seed, sweep_id)
