#!/usr/bin/env python
"""Benchmark of the MPP RJMCMC hot path (BASELINE.json metric: RJMCMC proposals/s and ms/image per B200).

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU sampler on the host cores

Workload (configs[2] of BASELINE.json): synthetic 2048x2048 DOTA-vehicle-like scene, ~2k oriented rectangles, hrcM
energy model (legacy setup + hierarchical combinator with the shipped weights / calibration).  One *step* = one image's
worth of sampling: `--sweeps` parallel colour sweeps, every 32-px window performing `--per-visit` local proposals per
sweep (defaults give ~2.4 M proposals; the reference's own budget for a 2048^2 image is 64 patches x 30 257 steps = 1.94 M).
N > 1: every rank samples its own scene (independent images -> weak scaling, no data-path collective).

`value`     : proposals/s with the maps already resident in HBM (CUDA events around K steps, max over ranks).
`e2e`       : the same metric through the public entry point api.sample_rjmcmc(ImageWMaps on the HOST, ...): pinned
              host maps -> H2D -> prefix sums -> naive initial configuration -> sweeps -> D2H of the final configuration
              as Rectangle objects, all inside the timed region.
`roofline`  : algorithmic bytes/proposal (recomputed with the run's K2 and acceptance) x proposals
              per k_sweep launch / mean launch duration, against the measured HBM peak of MEASURED_PEAKS.json.
`cpu_baseline`: the oracle's sequential sampler on the host cores, reference decomposition (256^2 patches, one process
              per core), bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ncu --set full capture of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one launch and the number of
# proposals that launch evaluated), written next to the raw CSV by tools/ncu_traffic.py; roofline.traffic is derived from it
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
KERNEL_NAMES = ["uniform_birth", "uniform_death", "data_birth", "data_death", "gaussian_translation", "data_translation",
                "gaussian_mark_transform", "data_mark_transform"]
METRIC = "rjmcmc_proposals_per_sec"
UNIT = "proposals/s"

# models_storage/mpp/mpp_hrcM/calibration.json + energy_combination_model.pkl (decoded)
CALIB_HRCM = dict(detection_threshold=0.6464646464646465,
                  coefs=(40.61517366849468, 37.691647645515616, 35.287160617991965),
                  intercepts=(-3.4763386136080006, -2.389875684851162, -3.8677248427633804),
                  min_area=23.553573615517713, max_area=166.58205129586045)
HRC = dict(weights_data=(0.8, 0.2), weights_prior=(0.7058823529411764, 0.058823529411764705, 0.23529411764705882),
           data_prior_weights=(0.5, 0.5), detection_threshold=0.0, bias=0.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--n-rect", type=int, default=0, help="candidate rectangles (0: 2600 per 2048^2, scaled by area)")
    ap.add_argument("--sweeps", type=int, default=6)
    ap.add_argument("--per-visit", type=int, default=96)
    ap.add_argument("--warps", type=int, default=8, help="warps per window (speculation depth) of the window sampler")
    ap.add_argument("--sampler", default="windows", choices=["windows", "cells"],
                    help="windows: mpp_run_windows (production); cells: mpp_run_sweeps (first-generation aligned cells)")
    ap.add_argument("--schedule", default="dataflow", choices=["dataflow", "colours", "multi"],
                    help="window sampler: one persistent dataflow kernel per call, or one launch per colour class")
    ap.add_argument("--stride", type=int, default=3)
    ap.add_argument("--temperature", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=4000, help="proposals per CPU worker in the cpu_baseline sample")
    ap.add_argument("--split", action="store_true",
                    help="BASELINE configs[3] alone: ONE scene split into row bands across the ranks (peer access over NVLink); strong scaling")
    ap.add_argument("--split-size", type=int, default=8192, help="side of the scene of the split-scene record (BASELINE configs[3])")
    ap.add_argument("--split-n-rect", type=int, default=0, help="candidate rectangles of that scene (0: 33000 per 8192^2 -> ~30k objects)")
    ap.add_argument("--no-split", action="store_true", help="skip the split-scene sub-record")
    ap.add_argument("--tiles", type=int, default=256, help="tiles of the tile-batch sub-record (BASELINE configs[4])")
    ap.add_argument("--tile-size", type=int, default=512)
    ap.add_argument("--no-tiles", action="store_true", help="skip the tile-batch sub-record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single-chain", action="store_true", help="skip the whole-scene single-chain CPU leg")
    return ap.parse_args()


def n_rect_for(args):
    return args.n_rect or int(round(2600 * (args.size * args.size) / (2048.0 * 2048.0)))


def workload_config(args):
    """The `config` object of BOTH arms (identical by construction, so that the driver's same-config check holds): what the
    workload is; how a run went is reported under `run`."""
    return {"workload": workload_name(args), "proposal_definition": "RJMCMC steps whose Delta-energy was evaluated",
            "l2": "inputs larger than L2 (mark maps 3 x %.0f MB per scene)" % (args.size * args.size * 32 * 4 / 1e6)}


def workload_name(args):
    return (f"synthetic {args.size}x{args.size} scene, {n_rect_for(args)} candidate rectangles (make_synth recipe), "
            f"hrcM energies, fixed T={args.temperature}")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ roofline
def bytes_per_proposal(k2: float, acceptance: float, h: int, w: int, p_kernel, window: int = 17) -> float:
    """B = 8*C + 20*(1+K2) + 16*p_add + 16*p_dataBD + 8*ceil(log2 H + log2 W)*p_dataBirth
    + 2*4*w^2*p_dataTrl + 128*p_dataTrf + 20*a   (C = 25 cells of the 5x5 block, K2 = mean objects in it).  p_kernel: the
    OBSERVED shares of the eight kernels among the evaluated proposals of the run (mpp_window_stats), not the reference
    mixture: empty windows only propose births, which moves most of the weight onto the two cheapest terms."""
    p = np.asarray(p_kernel, dtype=np.float64)
    p = p / max(p.sum(), 1e-300)
    p_add = 1.0 - (p[1] + p[3])
    p_data_bd = p[2] + p[3]
    return (8.0 * 25 + 20.0 * (1 + k2) + 16.0 * p_add + 16.0 * p_data_bd + 8.0 * math.ceil(math.log2(h) + math.log2(w)) * p[2]
            + 2 * 4.0 * window * window * p[5] + 128.0 * p[7] + 20.0 * acceptance)


def ncu_traffic():
    """(DRAM bytes per evaluated proposal, provenance) of the committed ncu capture, or (None, why)."""
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            t = json.load(f)
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) / float(t["proposals_in_launch"]), \
            {"file": os.path.relpath(NCU_TRAFFIC_FILE, ROOT), "source": t.get("source"), "kernel": t.get("kernel"),
             "dram_bytes_read": t["dram_bytes_read"], "dram_bytes_write": t["dram_bytes_write"],
             "proposals_in_launch": t["proposals_in_launch"], "l2_hit_pct": t.get("l2_hit_pct")}
    except (OSError, KeyError, ValueError) as e:
        return None, {"missing": f"{type(e).__name__}: {e}"}


def proposal_mix(ks, seconds):
    """Per-kernel tallies of the timed region (summed over ranks) -> the `proposal_mix` object of the bench line."""
    ev_e, ev_o = np.array(ks[0:8], dtype=np.float64), np.array(ks[8:16], dtype=np.float64)
    ac_e, ac_o = np.array(ks[16:24], dtype=np.float64), np.array(ks[24:32], dtype=np.float64)
    ev, ac = ev_e + ev_o, ac_e + ac_o
    tot = max(ev.sum(), 1.0)
    return {"kernels": KERNEL_NAMES, "evaluated": [int(v) for v in ev], "accepted": [int(v) for v in ac],
            "share": [float(v / tot) for v in ev], "reference_mixture": [1 / 18, 1 / 18, 1 / 9, 1 / 9, 1 / 9, 2 / 9, 1 / 9, 2 / 9],
            "evaluated_in_empty_windows": int(ev_e.sum()), "share_from_empty_windows": float(ev_e.sum() / tot),
            "evaluated_in_occupied_windows": int(ev_o.sum()),
            "acceptance_empty_windows": float(ac_e.sum() / max(ev_e.sum(), 1.0)), "acceptance_occupied_windows": float(ac_o.sum() / max(ev_o.sum(), 1.0)),
            "identity_accepted": int(ks[32]), "state_changing_accepted": int(ac.sum() - ks[32]),
            "window_visits": int(ks[33]), "visits_of_empty_windows": int(ks[34]),
            "value_occupied": float(ev_o.sum() / seconds), "value_occupied_unit": "proposals/s evaluated in windows that hold at least one object (the reference kernel mixture)",
            "state_changing_accepts_per_s": float((ac.sum() - ks[32]) / seconds),
            "note": "empty windows propose births only (evaluated ~96 per pass, almost all rejected at this temperature); "
                    "`value` counts them, `value_occupied` does not"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ reference arm
REF_STEPS_PER_PATCH = 30000 + 2 * 128 + 1  # mpp_hrcM rjmcmc_params: burn_in + num_samples * samples_interval (+1: stopping.py:42)


def scene_patches(objs, det, marks, limit=None):
    """The reference's inference tiling (mpp_model.py:231-249): 256^2 patches anchored at linspace(0, size - 256, ceil(size / 256)),
    i.e. 8 x 8 = 64 non-overlapping patches of a 2048^2 image.  Cropped on the device; only the patches are moved to the host."""
    from oracle import cpu_baseline as cb
    h, w = int(det.shape[0]), int(det.shape[1])
    ax = np.linspace(0, h - cb.PATCH, max(1, math.ceil(h / cb.PATCH)), dtype=int)
    ay = np.linspace(0, w - cb.PATCH, max(1, math.ceil(w / cb.PATCH)), dtype=int)
    patches = []
    for i in ax:
        for j in ay:
            if limit is not None and len(patches) >= limit:
                return patches
            i1, j1 = min(h, i + cb.PATCH), min(w, j + cb.PATCH)
            sel = (objs[:, 0] >= i) & (objs[:, 0] < i1) & (objs[:, 1] >= j) & (objs[:, 1] < j1)
            o = objs[sel].copy()
            o[:, 0] -= i
            o[:, 1] -= j
            patches.append((det[i:i1, j:j1].contiguous().cpu().numpy(),
                            [marks[k, i:i1, j:j1].contiguous().cpu().numpy() for k in range(3)], o))
    return patches


def make_pool(args, patches):
    from oracle import cpu_baseline as cb
    return cb.PatchPool(patches, "legacy", CALIB_HRCM, "hierarchical", HRC, workers=None, t0=args.temperature, alpha_t=1.0)


def cpu_image_legs(args, pool, patches, objs, det, marks):
    """BASELINE.md section 3 items 3-4 on the host cores: (i) ms per image of the reference's parallel mode -- every one of the
    image's 256^2 patches is sampled for a bounded number of steps, its measured rate is scaled to the shipped budget of 30 257
    steps per patch and the patches are scheduled on the cores; (ii) ONE sequential chain over the whole scene on one core."""
    from oracle import cpu_baseline as cb
    steps = max(100, args.cpu_steps // 8)
    t0 = time.perf_counter()
    res = pool.run_each(patches, steps, warm=30, seed=args.seed + 5000)
    wall = time.perf_counter() - t0
    rates = [r[0] / r[1] for r in res]
    t_img = cb.image_time_from_patch_rates(rates, REF_STEPS_PER_PATCH, pool.workers)
    out = {"ms_per_image": 1e3 * t_img, "patches": len(patches), "steps_per_patch_budget": REF_STEPS_PER_PATCH,
           "proposals_per_image": len(patches) * REF_STEPS_PER_PATCH, "cores": pool.workers,
           "proposals_per_s_whole_image": len(patches) * REF_STEPS_PER_PATCH / t_img,
           "per_patch_rate_min_median_max": [float(np.min(rates)), float(np.median(rates)), float(np.max(rates))],
           "how": f"all {len(patches)} patches of the scene sampled for {steps} steps each ({wall:.1f} s wall), per-patch rate scaled to the "
                  f"shipped budget of {REF_STEPS_PER_PATCH} steps, longest-first schedule on {pool.workers} cores (mpp_model.py:231-262)"}
    single = None
    if args.size <= 2048 and not args.no_single_chain:
        n = max(50, args.cpu_steps // 16)
        p, t, a, t_setup = cb.single_chain_whole_scene(det.cpu().numpy(), [marks[k].cpu().numpy() for k in range(3)], objs, "legacy",
                                                        CALIB_HRCM, "hierarchical", HRC, n_steps=n, n_warm=10, seed=args.seed, t0=args.temperature)
        single = {"value": p / t, "unit": UNIT, "cores": 1, "steps": n, "acceptance": a / max(1, p), "setup_s": t_setup,
                  "what": f"one sequential chain over the whole {args.size}^2 scene (RJMCMC.run without patch tiling, rjmcmc.py:172-181); "
                          "make_energies / make_kernels time reported as setup_s, not included"}
    return out, single


def run_reference(args):
    """The reference's CPU sampler (oracle port: the reference is pure Python and cannot travel to the GPU box) on all
    host cores, reference decomposition: one sequential chain per 256^2 patch per process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from mpp_cnn_rs_object_detection_b200 import synth
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    objs, det, marks = synth.make_scene_torch(args.seed, (args.size, args.size), n_rect_for(args), device)
    patches = scene_patches(objs, det, marks)
    pool = make_pool(args, patches)
    per_step = max(200, args.cpu_steps // 4)
    for _ in range(max(0, min(args.warmup, 1))):
        pool.run(per_step // 4, seed=args.seed + 1000)
    tot_p, tot_t = 0, 0.0
    for s in range(args.steps):
        p, t, _, _ = pool.run(per_step, seed=args.seed + 17 * s)
        tot_p += p
        tot_t += t
    image, single = cpu_image_legs(args, pool, patches, objs, det, marks)
    pool.close()
    value = tot_p / tot_t
    sample = f"{pool.workers} workers x {per_step} proposals per step on distinct 256^2 patches of the scene ({args.steps} steps)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "run": {"decomposition": "256x256 patches, one sequential chain per host core"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": pool.workers, "kind": "port", "sample": sample},
            "ms_per_image": image["ms_per_image"], "whole_image": image, "single_chain": single,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ product arm
def hrcm_spec():
    from mpp_cnn_rs_object_detection_b200.engine import ModelSpec
    return ModelSpec(setup="legacy", pos_threshold=CALIB_HRCM["detection_threshold"], remap_coefs=CALIB_HRCM["coefs"],
                     remap_intercepts=CALIB_HRCM["intercepts"], min_area=CALIB_HRCM["min_area"], max_area=CALIB_HRCM["max_area"],
                     combinator="hierarchical",
                     comb_w=list(HRC["weights_data"]) + list(HRC["weights_prior"]) + list(HRC["data_prior_weights"]) + [0.0],
                     comb_bias=HRC["bias"], comb_threshold=HRC["detection_threshold"])


def split_record(args, world, rank, device, dist):
    """BASELINE configs[3]: ONE dense scene (default 8192^2, ~30k objects) split into `world` row bands, one per GPU; every rank
    runs the persistent dataflow kernel over its band with peer access (NVLink P2P through CUDA IPC mappings) to the boundary
    cells of its neighbours -- no exchange step, no collective on the data path.  Band-local maps.  world == 1: the same scene on
    one GPU with mpp_run_windows (the strong-scaling reference).  Returns the sub-record (identical on every rank)."""
    import torch

    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg, synth
    from mpp_cnn_rs_object_detection_b200.engine import Engine

    h = w = args.split_size
    n_rect = args.split_n_rect or int(round(33000 * (h * w) / (8192.0 * 8192.0)))
    m0, m1 = mg.PeerSplitScene.map_rows(h, rank, world) if world > 1 else (0, h)
    t_gen = time.perf_counter()
    objs, det, marks = synth.make_scene_band_torch(args.seed, (h, w), n_rect, device, row0=m0, rows=m1 - m0)
    det_sum = float(np.sum(det))
    t_gen = time.perf_counter() - t_gen
    eng = Engine((h, w), device=device)
    det_band = torch.as_tensor(det[m0:m1]).to(device)
    if world > 1:
        eng.set_maps_band(det_band, marks, m0, det_sum)
    else:
        eng.set_maps(det_band, marks, det_sum=det_sum)
    eng.set_model(hrcm_spec())
    eng.set_kernels(intensity=max(1, len(objs)))
    uid = np.arange(len(objs))
    scene = mg.PeerSplitScene(eng, h, rank, world)
    sel = scene.select_initial(objs[:, :2]) if world > 1 else np.ones(len(objs), dtype=bool)
    eng.add_objects(objs[sel, :2], objs[sel, 2:5], uid=uid[sel])
    if world > 1:
        scene.attach_dist()

    def step(k):
        if world > 1:
            scene.run(args.sweeps, args.per_visit, args.warps, args.temperature, args.seed, sweep_offset=k * args.sweeps)
        else:
            eng.run_windows(args.sweeps, args.per_visit, args.warps, t0=args.temperature, seed=args.seed, sweep_offset=k * args.sweeps,
                            read_counters=False)

    steps, warm = max(2, min(args.steps, 5)), max(1, min(args.warmup, 3))
    for k in range(warm):
        step(k)
    torch.cuda.synchronize()
    eng.run_windows(0, args.per_visit, args.warps, t0=args.temperature)  # reads + resets the counters
    eng.window_stats()
    launches0 = eng.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(steps):
        step(warm + k)
    ev1.record()
    torch.cuda.synchronize()
    cnt = eng.run_windows(0, args.per_visit, args.warps, t0=args.temperature)
    n_end = len(eng)
    t = torch.tensor([float(cnt[4]), float(cnt[0]), float(cnt[1]), float(n_end), float(eng.launches - launches0)], dtype=torch.float64, device=device)
    tm = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    proposals, attempted, accepted, objects_end, launches = (float(v) for v in t.tolist())
    rec = {"workload": f"synthetic {h}x{w} dense scene, {len(objs)} objects at start (band generator, seed {args.seed}), hrcM energies, fixed T={args.temperature}, "
                       f"ONE scene split into {world} row band(s)",
           "value": proposals / (ms * 1e-3), "unit": UNIT, "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warm,
           "ms_per_step": ms / steps, "ms_per_image": ms / steps, "proposals_per_step": proposals / steps,
           "acceptance": accepted / max(1.0, proposals), "objects_start": len(objs), "objects_end": objects_end,
           "gpu_launches": int(launches), "launches_per_step_per_rank": launches / steps / world,
           "band_rows": [list(b) for b in mg.row_bands(h, world)], "map_rows_this_rank": [m0, m1],
           "map_bytes_this_rank": int(marks.numel() * 4 + det_band.numel() * 4),
           "transport": ("none (single GPU, mpp_run_windows dataflow schedule)" if world == 1 else
                         "NVLink P2P loads/stores of the neighbour bands' boundary cells inside k_windows_multi (CUDA IPC mappings); "
                         "completion stamps by st.release.sys into the neighbour's grid; no collective, no host synchronisation on the data path"),
           "timing": "CUDA events on each rank's stream around the timed steps, max over ranks", "scene_build_s": t_gen}
    if world > 1:
        scene.detach()
    eng.close()
    del marks, det_band
    torch.cuda.empty_cache()
    return rec


def tiles_record(args, world, rank, device, dist):
    """BASELINE configs[4]: a batch of independent 512x512 tiles sharded over the ranks (multi_gpu.shard_items), the position and
    shape U-Nets (PyTorch/cuDNN, random initialisation: no checkpoints travel) producing the detection and mark maps ON THE DEVICE,
    the sampler reading them in place (no host round trip, no .npy detour), all tiles of a rank in ONE persistent dataflow launch
    (api.sample_rjmcmc_tiles -> mpp_run_windows_batch).  Randomly initialised networks detect nothing, so a synthetic scene
    (make_synth recipe, ~160 vehicles per tile) is written into their output maps on the device before the sampler reads them.
    Timed per step, host wall clock with device synchronisation at the stage boundaries, max over ranks: U-Nets, set-up
    (prefix sums, naive initial configuration), sampling, read-back."""
    import torch

    import mpp_cnn_rs_object_detection_b200.api as api
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg, synth
    from mpp_cnn_rs_object_detection_b200.map_producers import MapProducer

    n_tiles, side = args.tiles, args.tile_size
    mine = mg.shard_items(n_tiles, world, rank)
    n_rect = int(round(2600 * side * side / (2048.0 * 2048.0)))
    objs_per_tile = [synth.make_objects(args.seed + 1000 + k, (side, side), n_rect) for k in mine]
    torch.manual_seed(1234)
    producer = MapProducer().to(device).eval()
    images = synth.render_tiles_torch(objs_per_tile, (side, side), device, seed=args.seed + rank)
    setup = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(
        CALIB_HRCM["detection_threshold"], list(CALIB_HRCM["coefs"]), list(CALIB_HRCM["intercepts"]), CALIB_HRCM["min_area"], CALIB_HRCM["max_area"]))
    comb = api.HierarchicalEnergyCombinator(np.array(HRC["weights_data"]), np.array(HRC["weights_prior"]), np.array(HRC["data_prior_weights"]),
                                            HRC["detection_threshold"], HRC["bias"])
    ncell = ((side + 31) // 32) ** 2
    budget = args.sweeps * ncell * args.per_visit
    params = dict(num_samples=1, energy_combinator=comb, init_config="naive", init_temperature=args.temperature, alpha_t=1.0, burn_in=budget - 3,
                  energy_setup=setup, samples_interval=1, target_temperature=0.0, proposals_per_visit=args.per_visit, warps_per_window=args.warps,
                  reuse_device_maps=False)
    rng = np.random.default_rng(args.seed + 300 + rank)
    chunk = 16  # tiles per U-Net forward pass

    def barrier():
        if world > 1:
            dist.barrier()

    def step():
        t0 = time.perf_counter()
        dets, markss = [], []
        for s0 in range(0, len(mine), chunk):
            d, m = producer.produce(images[s0:s0 + chunk])
            synth.inject_objects_torch(d, m, objs_per_tile[s0:s0 + chunk])
            dets.append(d); markss.append(m)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tiles = []
        for j, d in enumerate(dets):
            for q in range(d.shape[0]):
                tiles.append(api.ImageWMaps(name=f"tile_{rank}_{j * chunk + q}", shape=(side, side), image=None, detection_map=d[q],
                                            param_dist_maps=[markss[j][q, 0], markss[j][q, 1], markss[j][q, 2]], mappings=api.default_mappings(),
                                            param_names=["size", "ratio", "angle"]))
        res, st = api.sample_rjmcmc_tiles(tiles, rng, return_stats=True, **params)
        t2 = time.perf_counter()
        return {"unet_s": t1 - t0, "setup_s": st["setup_s"], "sample_s": st["sample_s"], "collect_s": st["collect_s"], "total_s": t2 - t0,
                "evaluated": st["evaluated"], "objects": sum(len(r) for r in res), "launches": st["sampler_launches"]}

    steps, warm = max(2, min(args.steps, 3)), 1
    for _ in range(warm):
        step()
    barrier()
    torch.cuda.synchronize()
    acc = None
    t_all = time.perf_counter()
    for _ in range(steps):
        r = step()
        acc = r if acc is None else {k: acc[k] + v for k, v in r.items()}
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t_all
    v = torch.tensor([t_all, acc["unet_s"], acc["setup_s"], acc["sample_s"], acc["collect_s"]], dtype=torch.float64, device=device)
    tot = torch.tensor([float(acc["evaluated"]), float(acc["objects"]), float(acc["launches"])], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_all, t_unet, t_setup, t_sample, t_collect = (float(x) for x in v.tolist())
    evaluated, objects, launches = (float(x) for x in tot.tolist())
    rec = {"workload": f"{n_tiles} synthetic {side}x{side} tiles (independent chains), ~{n_rect} candidate rectangles each, hrcM energies, fixed T={args.temperature}, "
                       f"maps produced on the device by randomly initialised position / shape U-Nets with the synthetic scene written into them",
           "n_gpus": world, "tiles": n_tiles, "tiles_per_rank": len(mine), "steps": steps, "scaling": "strong (the batch is fixed, tiles are sharded)",
           "tiles_per_s": n_tiles * steps / t_all, "value": evaluated / t_all, "unit": UNIT, "ms_per_batch": 1e3 * t_all / steps,
           "sampler_value": evaluated / t_sample, "sampler_value_note": "proposals/s over the sampling stage alone (one mpp_run_windows_batch launch per rank and step)",
           "stage_s_per_batch": {"unets": t_unet / steps, "setup": t_setup / steps, "sampling": t_sample / steps, "read_back": t_collect / steps},
           "unet_share": t_unet / t_all, "proposals_per_tile": evaluated / (n_tiles * steps), "objects_found_per_tile": objects / (n_tiles * steps),
           "sampler_launches_per_step_per_rank": launches / steps / world, "h2d_bytes_per_step": 0,
           "call": "map_producers.MapProducer.produce(images on device) -> api.sample_rjmcmc_tiles(ImageWMaps with DEVICE maps, init_config='naive')"}
    del producer, images
    torch.cuda.empty_cache()
    return rec


def run_split(args):
    """`bench.py --split`: the split-scene record alone, as the bench line (BASELINE configs[3])."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    clocks = ClockSampler(local)
    clocks.start()
    rec = split_record(args, world, rank, device, dist)
    clk = clocks.stop()
    line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": rec["steps"], "warmup": rec["warmup"],
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": rec["workload"], "sweeps_per_step": args.sweeps, "proposals_per_visit": args.per_visit,
                                            "warps_per_window": args.warps, "l2": "inputs larger than L2"},
            "ms_per_image": rec["ms_per_image"], "e2e": None, "gpu_launches": rec["gpu_launches"], "clocks": clk, "split": rec}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200(args):
    import torch
    import torch.distributed as dist

    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg, synth
    from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec, kernel_probabilities

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (product arm) needs a CUDA device: the MPP sampler has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # one process per GPU: keep each rank (and the pinned host buffers it allocates) on its GPU's NUMA node, so that the
    # uploads of all ranks do not cross the socket interconnect (single-GPU runs keep every core for the CPU baseline leg)
    numa_cpus = mg.bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    h = w = args.size
    objs, det, marks = synth.make_scene_torch(args.seed + rank, (h, w), n_rect_for(args), device)
    spec = hrcm_spec()
    eng = Engine((h, w), device=device)
    eng.set_maps(det, marks)
    eng.set_model(spec)
    eng.set_kernels(intensity=max(1, len(objs)))
    eng.add_objects(objs[:, :2], objs[:, 2:5])
    n0 = len(eng)

    def run(n_sweeps, seed_off, read_counters=False):
        if args.sampler == "windows" and args.schedule == "multi" and n_sweeps > 0:
            from mpp_cnn_rs_object_detection_b200.engine import run_windows_batch
            return run_windows_batch([eng], [args.seed + rank], n_sweeps, args.per_visit, n_warps=args.warps, t0=args.temperature,
                                     sweep_offset=seed_off * args.sweeps, read_counters=read_counters)
        if args.sampler == "windows":
            return eng.run_windows(n_sweeps, args.per_visit, args.warps, t0=args.temperature, alpha_t=1.0, seed=args.seed + rank,
                                   sweep_offset=seed_off * args.sweeps, read_counters=read_counters,
                                   schedule="dataflow" if args.schedule == "multi" else args.schedule)
        return eng.run_sweeps(n_sweeps, args.per_visit, args.stride, t0=args.temperature, alpha_t=1.0, seed=args.seed + rank,
                              sweep_offset=seed_off * args.sweeps, read_counters=read_counters)

    def step(seed_off):
        run(args.sweeps, seed_off)

    for s in range(args.warmup):
        step(s)
    run(0, 0, read_counters=True)  # reads + resets the counters
    eng.window_stats()             # ... and the per-kernel tallies
    launches0 = eng.launches
    clocks = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        step(args.warmup + s)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    clk = clocks.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    cnt = run(0, 0, read_counters=True)
    ks_local = eng.window_stats()
    ks_vec = (ks_local["evaluated_empty"] + ks_local["evaluated_occupied"] + ks_local["accepted_empty"] + ks_local["accepted_occupied"] +
              [ks_local["identity_accepted"], ks_local["visits"], ks_local["visits_empty"]])
    if world > 1:
        kt = torch.tensor(ks_vec, dtype=torch.float64, device=device)
        dist.all_reduce(kt, op=dist.ReduceOp.SUM)
        ks_vec = [float(v) for v in kt.tolist()]
    gpu_launches = eng.launches - launches0
    # a proposal = one RJMCMC step whose Delta-energy was evaluated (counter 4); attempts that end before the energy
    # evaluation (empty perturbations, moves leaving their window) are NOT counted
    proposals = sum_over_ranks(float(cnt[4]))
    attempted = sum_over_ranks(float(cnt[0]))
    accepted = sum_over_ranks(float(cnt[1]))
    n1 = len(eng)
    value = proposals / (ms * 1e-3)

    # ---- e2e: the public entry point api.sample_rjmcmc on HOST inputs (pinned maps), H2D and D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        import mpp_cnn_rs_object_detection_b200.api as api
        det_h = torch.empty(det.shape, dtype=torch.float32, pin_memory=True).copy_(det)
        marks_h = torch.empty(marks.shape, dtype=torch.float32, pin_memory=True).copy_(marks)
        image = api.ImageWMaps(name=f"synthetic_{rank}", shape=(h, w), image=None, detection_map=det_h,
                               param_dist_maps=[marks_h[0], marks_h[1], marks_h[2]], mappings=api.default_mappings(),
                               param_names=["size", "ratio", "angle"])
        setup = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(
            CALIB_HRCM["detection_threshold"], list(CALIB_HRCM["coefs"]), list(CALIB_HRCM["intercepts"]), CALIB_HRCM["min_area"],
            CALIB_HRCM["max_area"]))
        comb = api.HierarchicalEnergyCombinator(np.array(HRC["weights_data"]), np.array(HRC["weights_prior"]),
                                                np.array(HRC["data_prior_weights"]), HRC["detection_threshold"], HRC["bias"])
        ncell = ((h + 31) // 32) * ((w + 31) // 32)
        budget = args.sweeps * ncell * args.per_visit  # proposals per image, as in the device-resident steps
        e2e_rng = np.random.default_rng(args.seed + 100 + rank)
        tot = {"evaluated": 0, "launches": 0, "objects": 0}

        params = dict(num_samples=1, energy_combinator=comb, init_config="naive", init_temperature=args.temperature, alpha_t=1.0,
                      burn_in=budget - 3, energy_setup=setup, samples_interval=1, target_temperature=0.0,
                      proposals_per_visit=args.per_visit, warps_per_window=args.warps, return_stats=True)

        def e2e_run(n_images):
            # every image is uploaded from pinned host memory inside this call (upload of image i+1 overlaps the sampling of i)
            res = api.sample_rjmcmc_batch([image] * n_images, e2e_rng, **params)
            for rects, st in res:
                tot["evaluated"] += st["evaluated"]; tot["launches"] += st["launches"]; tot["objects"] = len(rects[0])

        def h2d_probe(nbytes=1 << 30, reps=4):
            # plain pinned cudaMemcpyAsync on every rank at once: the aggregate host-to-device rate this box gives N concurrent
            # uploaders, i.e. the ceiling of any end-to-end figure that uploads the maps
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
            dev.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            barrier()
            t_p = time.perf_counter()
            for _ in range(reps):
                dev.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            t_p = max_over_ranks(time.perf_counter() - t_p)
            barrier()
            return nbytes * reps * world / t_p / 1e9

        h2d_ceiling = h2d_probe()
        e2e_steps = max(2, min(args.steps, 16))  # images per timed batch; the first upload (pipeline fill) is inside the timed region
        e2e_run(2)
        tot.update(evaluated=0, launches=0)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        barrier()
        p2 = sum_over_ranks(float(tot["evaluated"]))
        e2e = {"value": p2 / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(det_h.numel() * 4 + marks_h.numel() * 4),
               "d2h_bytes_per_step": int(tot["objects"] * (4 + 8 + 24 + 4)), "steps": e2e_steps, "ms_per_image": 1e3 * t_e2e / e2e_steps,
               "objects_found": tot["objects"], "gpu_launches": int(tot["launches"]),
               "h2d_gbs": (det_h.numel() * 4 + marks_h.numel() * 4) * e2e_steps * world / t_e2e / 1e9,
               "h2d_ceiling_gbs": h2d_ceiling,
               "h2d_note": "h2d_gbs: map bytes uploaded by all ranks / e2e time; h2d_ceiling_gbs: aggregate rate of a plain pinned cudaMemcpyAsync "
                           "of 1 GiB x 4 issued by all ranks at once in this run (the platform's ceiling for N concurrent uploaders)",
               "numa_bound_cpus": (len(numa_cpus) if numa_cpus else None),
               "call": "api.sample_rjmcmc_batch([ImageWMaps with pinned host maps] x steps, init_config='naive', fixed T) -> List[Rectangle] per image",
               "timer": "host wall clock around the call; whole batch incl. pipeline fill; per image: H2D of its maps (overlapped with the "
                        "previous image's sampling) + prefix sums + naive init + sampler + D2H of the configuration (overlapped with the next image's sampling); max over ranks"}
        del det_h, marks_h, image

    # ---- roofline of the dominant kernel (k_sweep)
    ncell = ((h + 31) // 32) * ((w + 31) // 32)
    k2 = 25.0 * (0.5 * (n0 + n1)) / ncell
    acc = accepted / max(1.0, proposals)
    mix = proposal_mix(ks_vec, ms * 1e-3) if args.sampler == "windows" else None
    observed = np.array(mix["evaluated"], dtype=np.float64) if mix and sum(mix["evaluated"]) > 0 else kernel_probabilities()
    bpp = bytes_per_proposal(k2, acc, h, w, observed)
    bpp_reference_mixture = bytes_per_proposal(k2, acc, h, w, kernel_probabilities())
    peak, peak_src = measured_peak()
    if args.sampler == "windows" and args.schedule == "dataflow":
        sweep_launches = args.steps  # one persistent kernel per step
    else:
        sweep_launches = args.steps * args.sweeps * (9 if args.sampler == "windows" else args.stride * args.stride)
    achieved = bpp * (proposals / world) / (ms * 1e-3) / 1e9
    per_launch = proposals / world / sweep_launches
    traffic_pp, traffic_src = ncu_traffic()
    traffic = traffic_pp * per_launch if (traffic_pp and args.sampler == "windows" and args.schedule == "dataflow" and args.warps == 8) else None
    roofline = {"bound": "hbm", "kernel": ("k_windows_dataflow<float,%d>" % args.warps if args.schedule == "dataflow" else "k_sweep2<float,%d>" % args.warps) if args.sampler == "windows" else "k_sweep<float>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per evaluated proposal of the committed capture x proposals of this launch)",
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": bpp * per_launch, "peak_source": peak_src, "bytes_per_proposal": bpp,
                "bytes_per_proposal_basis": "SURVEY 8d formula with the OBSERVED kernel shares of this run (proposal_mix.share)",
                "bytes_per_proposal_reference_mixture": bpp_reference_mixture, "k2_objects_in_5x5": k2,
                "proposals_per_launch": proposals / world / sweep_launches, "us_per_launch": 1e3 * ms / sweep_launches,
                "note": "latency/parallelism-bound Markov chain: see DESIGN.md"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args),
            "run": {"sampler": args.sampler, "sweeps_per_step": args.sweeps,
                    "proposals_per_visit": args.per_visit, "warps_per_window": args.warps, "schedule": args.schedule if args.sampler == "windows" else "colours", "colour_stride": 3 if args.sampler == "windows" else args.stride,
                    "attempted_per_step": attempted / args.steps,
                    "proposal_definition": "RJMCMC steps whose Delta-energy was evaluated (empty perturbations and moves leaving their window are not counted)", "objects_start": n0, "objects_end": n1, "acceptance": acc,
                    "proposals_per_step": proposals / args.steps, "parallelism": f"{world} independent scene(s), one per GPU"},
            "ms_per_image": ms / args.steps, "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clk, "roofline": roofline,
            "proposal_mix": mix}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        patches = scene_patches(objs, det, marks)
        pool = make_pool(args, patches)
        p, t, a, _ = pool.run(args.cpu_steps, warm=100, seed=args.seed)
        image, single = cpu_image_legs(args, pool, patches, objs, det, marks)
        pool.close()
        line["cpu_baseline"] = {"value": p / t, "unit": UNIT, "cores": pool.workers, "kind": "port",
                                "sample": f"{pool.workers} workers x {args.cpu_steps} proposals, one 256^2 patch chain each (reference decomposition), "
                                          f"T={args.temperature}, acceptance {a / max(1, p):.3f}",
                                "per_core": p / t / pool.workers, "ms_per_image": image["ms_per_image"], "whole_image": image,
                                "single_chain": single}
    eng.close()
    del eng, det, marks
    Engine.drain_pool()
    torch.cuda.empty_cache()
    if not args.no_split and args.sampler == "windows":
        line["split"] = split_record(args, world, rank, device, dist)
    if not args.no_tiles and args.sampler == "windows":
        line["tiles"] = tiles_record(args, world, rank, device, dist)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.split:
        run_split(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
