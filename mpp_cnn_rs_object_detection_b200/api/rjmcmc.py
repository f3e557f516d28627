"""The Metropolis-Hastings-Green loop and its entry function: models/mpp/rjmcmc_sampler/{rjmcmc,sample_rjmcmc,
stopping}.py and utils/nms.py:68-109.

Three execution modes, all on the device:
  * RJMCMC.step()            one proposal per call through the Kernel objects (any Kernel subclass works);
  * RJMCMC.run()             with the built-in kernels and StopOnMaxIter: the whole sequential chain in one kernel
                             (mpp_run_chain: the reference's algorithm, global kernels, Philox randomness);
  * sample_rjmcmc(..., sampler='parallel')  the colour-sweep sampler (mpp_run_sweeps): same target distribution and
                             the same proposal budget / temperature trajectory, cells sampled concurrently."""
from __future__ import annotations

import logging
import time
import warnings
from dataclasses import dataclass
from typing import Callable, Dict, List, Tuple, Union

import numpy as np

from .custom_types import EnergyCombinationModel, ImageWMaps, RJMCMCStateSummary
from .energy_point_set import EPointsSet
from .energy_setups import EnergySetup
from .kernels import DeviceKernel, Kernel, make_kernels
from .shapes import Rectangle

EPS = 1e-16  # rjmcmc.py:15


# ---------------------------------------------------------------------------------------------- stopping.py
class StoppingCondition:
    def do_stop(self, states: List[RJMCMCStateSummary]) -> bool:
        raise NotImplementedError

    def print(self, states: List[RJMCMCStateSummary]) -> str:
        return ""


class StopOnMaxIter(StoppingCondition):  # stopping.py:37-45
    def __init__(self, max_iter: int):
        self.max_iter = max_iter

    def do_stop(self, states: List[RJMCMCStateSummary]) -> bool:
        return states[-1].iter >= self.max_iter

    def print(self, states: List[RJMCMCStateSummary]) -> str:
        return f"{states[-1].iter} < {self.max_iter}"


# ---------------------------------------------------------------------------------------------- rjmcmc.py
class RJMCMCTimer:
    """Per-stage wall-clock of the step-by-step chain, same interface and stage names as the reference (rjmcmc.py:18-48):
    `timings` maps 'sample_kernel', 'sample_perturbation', 'compute_energy', 'compute_alpha', 'apply_perturbation', 'log' to
    one duration per step, plus 'total' and 'n_points'.  Every stage is a call into the device library here, so a duration is
    launch + synchronisation latency, not CPU arithmetic.  The device-resident chains (the sequential chain run as one kernel
    and the window sampler) have no host stages: they add one 'total' entry per launch and report their counters through
    sample_rjmcmc(return_stats=True) / mpp_window_stats."""

    def __init__(self):
        self.last_tick = None
        self.timings: Dict[str, List[float]] = {"total": [], "n_points": []}
        self.start_tick = None

    def start_step(self):
        self.start_tick = time.perf_counter()
        self.last_tick = self.start_tick

    def checkpoint(self, key):
        now = time.perf_counter()
        self.timings.setdefault(key, []).append(now - self.last_tick)
        self.last_tick = now

    def end_step(self, n_points):
        self.timings["total"].append(time.perf_counter() - self.start_tick)
        self.timings["n_points"].append(n_points)

    def show_results(self):
        points_number = np.array(self.timings["n_points"])
        for k, l in self.timings.items():
            if k != "n_points":
                l = np.array(l)
                per_point = l[:len(points_number)][points_number[:len(l)] > 0] / points_number[:len(l)][points_number[:len(l)] > 0] if len(l) else l
                print(f"{k:20}: {np.mean(l) if len(l) else 0.0:.2e} s | {np.mean(per_point) if len(per_point) else 0.0:.2e} s/point")


@dataclass
class RJMCMC:
    t0: float
    kernels: List[Kernel]
    p_kernels: List[float]
    initial_state: EPointsSet
    stopping_condition: StoppingCondition
    rng: np.random.Generator
    energy_combinator: EnergyCombinationModel = None
    t_target: float = 0
    sampling_rule: Callable[[int], bool] = None
    do_annealing = True
    alpha_t: float = None
    verbose: int = 0

    def __post_init__(self):
        assert len(self.kernels) == len(self.p_kernels)
        assert (not self.do_annealing) or (self.alpha_t is not None)
        assert self.t0 >= self.t_target
        self._temp: float = self.t0
        self._iter: int = 0
        self._state_log: List[EPointsSet] = [self.initial_state]
        self._state_summaries: List[RJMCMCStateSummary] = [RJMCMCStateSummary(n_points=len(self.initial_state), iter=self._iter)]
        self._timer = RJMCMCTimer()

    def get_timings(self) -> RJMCMCTimer:
        """rjmcmc.py:183-184."""
        return self._timer

    def get_state_log(self):
        """rjmcmc.py:186-187."""
        return self._state_log

    def step(self, return_state=False):
        """One Metropolis-Hastings-Green step (rjmcmc.py:83-164)."""
        if self.stopping_condition.do_stop(self._state_summaries):
            raise StopIteration
        self._timer.start_step()
        k1: Kernel = self.kernels[int(self.rng.choice(len(self.kernels), p=self.p_kernels))]
        self._timer.checkpoint("sample_kernel")
        x0 = self._state_log[-1]
        u1 = k1.sample_perturbation(x0.points, self.rng)
        self._timer.checkpoint("sample_perturbation")
        energy_x0 = self._state_summaries[-1].energy
        if energy_x0 is None:
            energy_x0 = x0.total_energy()  # raw sum on the first iteration (rjmcmc.py:96-98)
        energy_delta = x0.energy_delta(u1, energy_combinator=self.energy_combinator)
        energy_x1 = energy_x0 + energy_delta
        self._timer.checkpoint("compute_energy")
        log_alpha_1 = (-energy_delta / self._temp) + np.log(k1.backward_probability(x0.points, u1) + EPS) \
            - np.log(k1.forward_probability(x0.points, u1) + EPS)
        accepted = bool(np.log(self.rng.random() + EPS) < log_alpha_1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            alpha_1 = float(np.exp(log_alpha_1))
        self._timer.checkpoint("compute_alpha")
        x1 = x0.apply_perturbation(u1, inplace=True) if accepted else x0
        self._timer.checkpoint("apply_perturbation")
        summary = RJMCMCStateSummary(iter=self._iter, temperature=self._temp, energy=energy_x1 if accepted else energy_x0,
                                     n_points=len(x1), kernel=k1.__class__, move_accepted=accepted, alpha=alpha_1,
                                     initial_energy=energy_x0, proposed_energy=energy_x1)
        self._state_summaries.append(summary)
        if self.sampling_rule is not None and self.sampling_rule(self._iter):
            self._state_log.append(x1.copy())
        else:
            self._state_log[0] = x1
        self._iter += 1
        if self.do_annealing and self._temp > self.t_target:
            self._temp *= self.alpha_t
        self._timer.checkpoint("log")
        self._timer.end_step(n_points=len(x1))
        if return_state:
            return summary, x1.copy()
        return summary

    def __iter__(self):
        return self

    def __next__(self):
        return self.step()

    # -- whole-chain fast path -------------------------------------------------------------------------
    def _device_chain_possible(self) -> bool:
        if not isinstance(self.stopping_condition, StopOnMaxIter) or len(self.kernels) != 8:
            return False
        if not all(isinstance(k, DeviceKernel) and k.KERNEL_ID == i for i, k in enumerate(self.kernels)):
            return False
        if len({id(k._kset) for k in self.kernels}) != 1:
            return False
        return self.energy_combinator is None or hasattr(self.energy_combinator, "device_params")

    def _run_on_device(self):
        x = self._state_log[-1]
        st = x._state
        kset = self.kernels[0]._kset
        kset.p_kernels = np.asarray(self.p_kernels, dtype=np.float64)
        kset.bind(x.points)
        st.use_combinator(self.energy_combinator)
        total = self.stopping_condition.max_iter + 1 - self._iter  # StopOnMaxIter runs max_iter + 1 steps (stopping.py:42)
        seed = int(self.rng.integers(0, 2 ** 62))
        alpha = self.alpha_t if self.do_annealing else 1.0
        done = 0
        while done < total:
            # run up to (and including) the next step at which sampling_rule asks for a snapshot
            seg_end = total
            if self.sampling_rule is not None:
                for s in range(done, total):
                    if self.sampling_rule(self._iter + (s - done)):
                        seg_end = s + 1
                        break
            n = seg_end - done
            _, trace = st.engine.run_chain(n, t0=self._temp, alpha_t=alpha, t_target=self.t_target, seed=seed, step_offset=self._iter,
                                           trace=True)
            for row in trace:
                self._state_summaries.append(RJMCMCStateSummary(
                    iter=self._iter, temperature=float(row["temperature"]), n_points=int(row["n_after"]), move_accepted=bool(row["accepted"]),
                    alpha=float(np.exp(min(row["log_alpha"], 700.0)))))
                self._iter += 1
                if self.do_annealing and self._temp > self.t_target:
                    self._temp *= self.alpha_t
            st.refresh_from_device()
            x.energy_graph._members = dict.fromkeys(st.handle_of)
            done = seg_end
            if self.sampling_rule is not None and self.sampling_rule(self._iter - 1):
                self._state_log.append(x.copy())

    def run(self, show_timing=False) -> Tuple[Union[List[EPointsSet], EPointsSet], List[RJMCMCStateSummary]]:
        if self._device_chain_possible() and self.verbose == 0:
            self._timer.start_step()
            self._run_on_device()
            self._timer.end_step(n_points=len(self._state_log[-1]))
        else:
            for _ in self.__iter__():
                pass
        if show_timing:
            print("Timings--------------------------------------------------------------------------")
            self._timer.show_results()
        return self._state_log, self._state_summaries


# ---------------------------------------------------------------------------------------------- utils/nms.py:68-109
def nms_distance(centers, confidence_score, threshold, return_index=False):
    """Greedy distance NMS (highest score first; everything within `threshold` of a picked centre is dropped).
    Host-side helper kept for API compatibility; sample_rjmcmc's 'naive' init runs mpp_naive_init on the device."""
    if len(centers) == 0:
        return ([], [], []) if return_index else ([], [])
    centers = np.asarray(centers)
    score = np.asarray(confidence_score)
    order = np.argsort(score)
    picked = []
    while order.size > 0:
        idx = order[-1]
        picked.append(idx)
        d = np.linalg.norm(centers[idx] - centers[order[:-1]], axis=-1)
        order = order[:-1][d > threshold]
    pc, ps = [centers[i] for i in picked], [confidence_score[i] for i in picked]
    return (pc, ps, picked) if return_index else (pc, ps)


def naive_detection(image_data: ImageWMaps, detection_threshold: float, energy_setup: EnergySetup = None) -> List[Rectangle]:
    """Threshold -> greedy 6-px NMS -> argmax marks (sample_rjmcmc.py:23-35), on the device (mpp_naive_init)."""
    from .energy_setups import LegacyEnergiesCalibration, LegacyEnergySetup
    if energy_setup is None:
        energy_setup = LegacyEnergySetup(energy_calibration=LegacyEnergiesCalibration(detection_threshold, [1, 1, 1], [0, 0, 0], 0, 1e30))
    unit, pair = energy_setup.make_energies(image_data)
    pts = EPointsSet([], image_data.shape, unit, pair)
    pts._state.engine.naive_init(float(detection_threshold), 6.0)
    pts._state.refresh_from_device()
    return list(pts._state.objects())


def plan_sweeps(shape, seed: int, budget: int, proposals_per_visit: int, sweep_offset: int = 0):
    """(proposals per visit, number of sweeps, mean proposals per sweep) of the window sampler for a budget of `budget` RJMCMC
    steps (max_iter + 1 of the reference's loop, stopping.py:42): counts the windows of every sweep's shifted grid (the host twin
    of mpp_window_grid gives the offsets) instead of assuming the aligned grid, lowers the proposals per visit when one sweep at
    the requested value would already exceed the budget, and stops at the first sweep that reaches it."""
    from ..multi_gpu import grid_offset
    h, w = int(shape[0]), int(shape[1])

    def windows(s):
        ox, oy = grid_offset(seed, sweep_offset + s)
        return ((h + ox + 31) // 32) * ((w + oy + 31) // 32)

    mean_windows = float(np.mean([windows(s) for s in range(16)]))
    pv = int(max(1, min(proposals_per_visit, int(np.ceil(budget / mean_windows)))))
    done, n = 0, 0
    while done < budget:
        done += windows(n) * pv
        n += 1
    return pv, n, done / n


# ---------------------------------------------------------------------------------------------- sample_rjmcmc.py
def sample_rjmcmc(image_data: ImageWMaps, rng: np.random.Generator, num_samples: int, energy_combinator: EnergyCombinationModel,
                  init_config: Union[str, List[Rectangle], None], init_temperature: float, alpha_t: Union[float, str], burn_in: int,
                  energy_setup: EnergySetup, samples_interval: int, target_temperature: float, verbose: int = 0,
                  iter_multiplier: float = None, use_split_merge: bool = False, sampler: str = "parallel",
                  proposals_per_visit: int = 96, warps_per_window: int = 8, precision: str = "fp32",
                  reuse_device_maps: bool = True, return_stats: bool = False, _device_maps=None, _defer: bool = False):
    """Drop-in for sample_rjmcmc (sample_rjmcmc.py:38-102): returns a list of `num_samples` PointsSet of Rectangle.

    sampler='parallel'   window sampler (mpp_run_windows): max_iter + 1 proposals in total, spread over
                         ceil(.. / (windows * proposals_per_visit)) sweeps; the temperature follows the reference's geometric
                         schedule as a function of the number of proposals made (one multiplication by
                         alpha_t ** proposals_per_sweep per sweep).
    sampler='sequential' the reference's one-proposal-at-a-time chain, run on the device.
    reuse_device_maps    False: always upload the maps (no cache keyed on the host arrays).
    return_stats         True: returns (result, dict of device counters) instead of result.
    _defer               (internal, sample_rjmcmc_batch) parallel sampler, single sample: queue the chain and return a function
                         that waits for it and builds the result, so that the caller can overlap host work with the sampling."""
    if use_split_merge and sampler != "sequential":
        raise NotImplementedError("the optional split / merge kernels run through the step-by-step chain: pass sampler='sequential'")
    unit_energies, pair_energies = energy_setup.make_energies(image_data)
    points = EPointsSet(points=[], support_shape=image_data.shape, unit_energies_constructors=unit_energies,
                        pair_energies_constructors=pair_energies, precision=precision, reuse_device_maps=reuse_device_maps,
                        _device_maps=_device_maps)
    st = points._state
    if isinstance(init_config, str) and init_config == "gt":
        st.add_many(image_data.gt_config)
    elif isinstance(init_config, str) and init_config == "naive":
        st.engine.naive_init(float(energy_setup.detection_threshold), 6.0)
        if sampler != "parallel":
            st.refresh_from_device()  # the parallel sampler only needs the count; the host mirror is rebuilt once at the end
    elif init_config is not None:
        st.add_many(list(init_config))
    points.energy_graph._members = dict.fromkeys(st.handle_of)

    if iter_multiplier is not None:  # sample_rjmcmc.py:58-61
        burn_in = burn_in * iter_multiplier
        samples_interval = samples_interval * iter_multiplier
        alpha_t = np.power(alpha_t, 1 / iter_multiplier)
    if isinstance(alpha_t, str) and alpha_t == "auto":  # :63-66
        alpha_t = np.power(target_temperature / init_temperature, 1 / burn_in)
        target_temperature = 0
    burn_in, samples_interval = int(burn_in), int(samples_interval)
    intensity = max(1, len(st.engine))  # :68
    kernels, p_kernels = make_kernels(image_data, intensity=intensity, rng=rng, use_split_merge=use_split_merge)
    max_iter = burn_in + (num_samples + 1) * samples_interval  # :78
    start = time.perf_counter()
    if sampler == "sequential":
        chain = RJMCMC(t0=init_temperature, t_target=target_temperature, alpha_t=alpha_t, kernels=kernels, p_kernels=p_kernels,
                       initial_state=points, energy_combinator=energy_combinator, stopping_condition=StopOnMaxIter(max_iter), rng=rng,
                       sampling_rule=lambda step: step >= burn_in and step % samples_interval == 0, verbose=verbose)
        states, _ = chain.run()
        result = [states[-1].points] if num_samples == 1 else [s.points for s in states[-num_samples:]]
    elif sampler == "parallel":
        import torch
        torch.cuda.nvtx.range_push("mpp.sample_rjmcmc.windows")  # NVTX: visible in Nsight Systems timelines
        kernels[0]._kset.bind(points.points)
        st.use_combinator(energy_combinator)
        eng = st.engine
        stats = {"proposals": 0, "accepted": 0, "births": 0, "deaths": 0, "evaluated": 0, "sweeps": 0}
        seed = int(rng.integers(0, 2 ** 62))
        # budget -> sweeps: a sweep visits every window of the shifted grid once ((H + ox + 31) // 32 x (W + oy + 31) // 32 of them,
        # up to one row and one column more than the aligned grid), `proposals_per_visit` proposals each; on a small budget the
        # proposals per visit are lowered so that one sweep does not exceed it
        proposals_per_visit, total_sweeps, per_sweep = plan_sweeps(image_data.shape, seed, max_iter + 1, int(proposals_per_visit))
        # snapshot steps of the reference's sampling_rule, expressed in sweeps
        # num_samples == 1: the reference returns its last snapshot, at most samples_interval - 1 steps before the end of the
        # chain; the final state is returned here instead (no intermediate read-back)
        snap_steps = [s for s in range(burn_in, max_iter + 1) if s % samples_interval == 0] if (samples_interval > 0 and num_samples > 1) else []
        snap_sweeps = sorted({min(total_sweeps, max(1, int(np.ceil((s + 1) / per_sweep)))) for s in snap_steps})
        alpha_sweep = float(np.power(alpha_t, per_sweep))
        temp, done, states = float(init_temperature), 0, []
        for stop in snap_sweeps + ([total_sweeps] if (not snap_sweeps or snap_sweeps[-1] < total_sweeps) else []):
            n = stop - done
            if n > 0:
                eng.run_windows(n, proposals_per_visit, warps_per_window, t0=temp, alpha_t=alpha_sweep,
                                t_target=float(target_temperature), seed=seed, sweep_offset=done, read_counters=False)
                stats["sweeps"] += n
                for _ in range(n):
                    if temp > target_temperature:
                        temp *= alpha_sweep
                done = stop
            if stop in snap_sweeps:
                st.refresh_from_device()
                points.energy_graph._members = dict.fromkeys(st.handle_of)
                states.append(points.copy())

        torch.cuda.nvtx.range_pop()

        def finish():
            if return_stats:
                c = eng.run_windows(0, proposals_per_visit, warps_per_window, t0=max(temp, 1e-30))
                stats.update(proposals=c[0], accepted=c[1], births=c[2], deaths=c[3], evaluated=c[4], launches=eng.launches)
            st.refresh_from_device()
            points.energy_graph._members = dict.fromkeys(st.handle_of)
            sampled = states if states else [points]
            res = [sampled[-1].points] if num_samples == 1 else [s.points for s in sampled[-num_samples:]]
            return (res, stats) if return_stats else res

        if _defer:
            return finish
        result = finish()
    else:
        raise ValueError(f"sampler must be 'parallel' or 'sequential', got {sampler!r}")
    end = time.perf_counter()
    if return_stats and sampler != "parallel":
        result = (result, {"proposals": max_iter + 1, "evaluated": max_iter + 1, "launches": st.engine.launches})
    logging.info(f"rjmcmc on image {image_data.name} ran in {end - start:.2f}s ({(end - start) / max(1, max_iter):.1e}s/iter) "
                 f"(int. {intensity} | iter {max_iter} | num_samples {num_samples} | {sampler})")
    return result


def sample_rjmcmc_batch(images, rng: np.random.Generator, with_scores: bool = False, **params):
    """sample_rjmcmc over a sequence of images (the role of `_map_to_images(partial(sample_rjmcmc, ...), images)`,
    models/mpp/train_energy_combination/train_utils.py:11-18 and mpp_model.py:250-264) with the host-to-device upload of
    image i+1 and the host-side read-back of image i-1 overlapped with the sampling of image i (second CUDA stream,
    recycled device buffers).  `params` are
    sample_rjmcmc's keyword arguments.  Device state is released image by image, so the result is detached: per image a
    list (one entry per sample) of lists of Rectangle; with_scores=True returns (rectangles, Papangelou scores of the last
    sample: exp(+Delta E of removal), mpp_model.py:296-304) per image; return_stats=True appends the counters dict."""
    import torch

    from .device_state import build_layout, device_maps
    images = list(images)
    if not images:
        return []
    energy_setup = params["energy_setup"]
    dev = torch.device("cuda", torch.cuda.current_device())
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(device=dev)

    def upload(img, first=False):
        unit, pair = energy_setup.make_energies(img)
        layout = build_layout(unit, pair)
        if first:  # later hand-outs recycle buffers whose last reader finished before a host synchronisation (see below)
            copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            dm = device_maps(layout.det, layout.marks, dev, reuse=False)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dm, ev

    def collect(out, dm):
        stats = None
        if callable(out):
            out = out()  # waits for the chain, reads the configuration back
        if params.get("return_stats"):
            out, stats = out
        item = [list(ps) for ps in out]
        if with_scores:
            st = out[-1]._state
            st.use_combinator(params.get("energy_combinator"))
            objs = st.objects()
            rec = np.concatenate([st.proposal_record(u, None) for u in objs]) if objs else None
            item = (item, np.exp(st.engine.delta_batch(rec)) if objs else np.zeros(0))
        return (item, stats) if stats is not None else item

    # three images are in flight: maps of i+1 crossing PCIe, chain i on the SMs, configuration of i-1 being read back and turned
    # into Rectangles on the host.  Device contexts and upload buffers go back to their pools only after collect(), i.e. after
    # a host synchronisation that follows their last kernel.
    defer = (params.get("sampler", "parallel") == "parallel" and params.get("num_samples", 1) == 1 and not with_scores)
    results, pending = [], None
    nxt = upload(images[0], first=True)
    for i, img in enumerate(images):
        dm, ev = nxt
        nxt = upload(images[i + 1]) if i + 1 < len(images) else None
        main.wait_event(ev)
        out = sample_rjmcmc(image_data=img, rng=rng, _device_maps=dm, _defer=defer, **params)
        if pending is not None:
            results.append(collect(*pending))
        pending = (out, dm)
        if not defer:
            results.append(collect(*pending))
            pending = None
        del out, dm
    if pending is not None:
        results.append(collect(*pending))
    return results


def sample_rjmcmc_tiles(images, rng: np.random.Generator, return_stats: bool = False, n_streams: int = None, **params):
    """Independent chains on a batch of small tiles (BASELINE configs[4]: e.g. 256 tiles of 512x512; the reference maps such
    patches over a process pool, mpp_model.py:231-264).  All tiles of one shape are sampled by ONE persistent dataflow launch
    (mpp_run_windows_batch): one 512^2 tile exposes only ~36 concurrently active windows, a batch fills the GPU like one large
    scene.  The maps of a tile may be device tensors (e.g. straight from map_producers.MapProducer): they are used in place.
    Single-sample, parallel sampler only; `params` as sample_rjmcmc (init_config 'naive' / 'gt' / list / None).  Returns one list
    of Rectangle per tile (and, with return_stats=True, a dict of counters and stage times).  `n_streams` is accepted for
    compatibility with the first-generation implementation (one stream and one launch per tile) and ignored."""
    import torch

    from ..engine import run_windows_batch
    images = list(images)
    if params.get("num_samples", 1) != 1 or params.get("sampler", "parallel") != "parallel":
        raise ValueError("sample_rjmcmc_tiles runs one parallel-sampler chain per tile (num_samples=1)")
    if params.get("use_split_merge"):
        raise NotImplementedError("the optional split / merge kernels run through the step-by-step chain (sample_rjmcmc(sampler='sequential'))")
    if params.get("precision", "fp32") != "fp32":
        raise NotImplementedError("the window sampler is float32 only")
    energy_setup, comb = params["energy_setup"], params["energy_combinator"]
    init_config = params.get("init_config", "naive")
    burn_in, interval = params["burn_in"], params["samples_interval"]
    alpha_t, t0, t_target = params["alpha_t"], params["init_temperature"], params["target_temperature"]
    mult = params.get("iter_multiplier")
    if mult is not None:
        burn_in, interval, alpha_t = burn_in * mult, interval * mult, np.power(alpha_t, 1 / mult)
    if isinstance(alpha_t, str) and alpha_t == "auto":
        alpha_t, t_target = np.power(t_target / t0, 1 / burn_in), 0
    pv, nw = int(params.get("proposals_per_visit", 96)), int(params.get("warps_per_window", 8))
    max_iter = int(burn_in) + 2 * int(interval)
    t_start = time.perf_counter()
    states = []
    for img in images:  # phase 1: index the maps, initial configuration
        unit, pair = energy_setup.make_energies(img)
        pts = EPointsSet([], img.shape, unit, pair, reuse_device_maps=params.get("reuse_device_maps", True))
        st = pts._state
        if isinstance(init_config, str) and init_config == "gt":
            st.add_many(img.gt_config)
        elif isinstance(init_config, str) and init_config == "naive":
            st.engine.naive_init(float(energy_setup.detection_threshold), 6.0)
        elif init_config is not None:
            st.add_many(list(init_config))
        n0 = len(st.engine)
        st.use_kernels(max(1, n0), kernel_probabilities_default())
        st.use_combinator(comb)
        states.append(pts)
    t_setup = time.perf_counter()
    seeds = [int(rng.integers(0, 2 ** 62)) for _ in images]
    by_shape = {}
    for k, img in enumerate(images):
        by_shape.setdefault(tuple(img.shape[:2]), []).append(k)
    totals = np.zeros(8, dtype=np.int64)
    n_launches = 0
    for shape, idx in by_shape.items():  # phase 2: one launch per shape
        pv_eff, n_sweeps, per_sweep = plan_sweeps(shape, seeds[idx[0]], max_iter + 1, pv)
        cnt = run_windows_batch([states[k]._state.engine for k in idx], [seeds[k] for k in idx], n_sweeps, pv_eff, n_warps=nw, t0=float(t0),
                                alpha_t=float(np.power(alpha_t, per_sweep)), t_target=float(t_target), grid_seed=seeds[idx[0]],
                                read_counters=return_stats)
        n_launches += 1
        if cnt is not None:
            totals += np.array(cnt, dtype=np.int64)
    if return_stats:
        torch.cuda.synchronize()
    t_sample = time.perf_counter()
    out = []
    for pts in states:  # phase 3: collect
        pts._state.refresh_from_device()
        out.append(list(pts._state.objects()))
    t_end = time.perf_counter()
    if return_stats:
        return out, {"proposals": int(totals[0]), "accepted": int(totals[1]), "births": int(totals[2]), "deaths": int(totals[3]),
                     "evaluated": int(totals[4]), "sampler_launches": n_launches, "setup_s": t_setup - t_start,
                     "sample_s": t_sample - t_setup, "collect_s": t_end - t_sample}
    return out


def kernel_probabilities_default():
    from ..engine import kernel_probabilities
    return kernel_probabilities()
