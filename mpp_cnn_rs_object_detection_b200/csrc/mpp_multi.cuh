// K6, multi-scene / split-scene schedule: ONE persistent dataflow kernel over the window visits of
//   * a batch of independent scenes of equal shape (tile batches: mpp_model.py:231-264 maps 256^2 patches over a process
//     pool; here all tiles of a rank share one launch so that a batch fills the GPU like one large scene), or
//   * the band of ONE scene that this GPU owns, the other bands being sampled at the same time by the same kernel on the
//     neighbour GPUs.  Nothing is packed or exchanged between phases: a window next to a band boundary stages the
//     neighbour's boundary cells with peer loads over NVLink, commits into them with peer stores, and window completion is
//     published into the neighbour's completion grid with a system-scope release store, so that the dependency wait of
//     the dataflow schedule simply extends across the GPUs.
//
// Visits are numbered in (sweep, colour, scene, window) order and claimed in that order from a counter.  A visit starts
// when every earlier visit within 64 px of its window has completed (see k_windows_dataflow); completion grids hold
// monotone sweep stamps that are never reset, so dependencies also hold across launches (a neighbour GPU may be a launch
// ahead).  Dependencies always precede in the global order and every rank claims in that order, so the globally first
// unfinished visit can always run: no deadlock, provided each rank's grid is resident (persistent grid <= occupancy).
// The chain is the chain of mpp_run_windows, bit for bit (uids, random streams and grid offsets depend only on
// (seed, sweep, window)).
#pragma once
#include "mpp_sweep2.cuh"

template <typename R>
struct SceneDev {
    Ctx<R> c;
    unsigned long long seed;      // random streams of this scene
    int *done;                    // [2][dg][dg] completion stamps of this scene, by stamp parity
    int *done_up, *done_down;     // the neighbour ranks' grids (split scene), or NULL
    int notify_lo, notify_hi;     // a visit whose window starts above pixel row notify_lo / ends below notify_hi also stamps
                                  // done_up / done_down (it can be a dependency of a window the neighbour owns)
};

struct MultiPlan {
    int n_sweeps, n_scenes, total_tasks, dg;
    int stamp0;             // stamp of the first sweep of this launch = sweeps run before on these contexts + 1
    const int *ox, *oy;     // [n_sweeps + 1] grid offsets; entry 0 is the sweep BEFORE this launch
    const int *wi_lo, *wi_hi;  // [n_sweeps] window rows of this rank: those whose first pixel row lies in its band
    const int *task_base;   // [9 * n_sweeps + 1] first task of every (sweep, colour)
    const float *temp;      // [n_sweeps]
    int *next_task;
};

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_sys(const int *p) {
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename R, int NW, bool DBG, bool SPLIT>
__global__ void __launch_bounds__(32 * NW, MPP_DF_MIN_BLOCKS) k_windows_multi(const __grid_constant__ SceneDev<R> scene0,
                                                                            const SceneDev<R> *__restrict__ scenes, MultiPlan plan, int per_visit,
                                                                            uint64_t sweep_offset, float *dbg_maxdiff) {
    extern __shared__ __align__(16) unsigned char smem[];
    WinState<R> &w = *reinterpret_cast<WinState<R> *>(smem);
    constexpr size_t WS = (sizeof(WinState<R>) + 15) & ~(size_t)15;
    R *scratch = reinterpret_cast<R *>(smem + WS);
    constexpr size_t SC = ((size_t)NW * W2_SCRATCH * sizeof(R) + 15) & ~(size_t)15;
    SceneDev<R> &sc = *reinterpret_cast<SceneDev<R> *>(smem + WS + SC);  // this visit's scene (shared-memory copy)
    __shared__ int s_task;
    // the plan's small arrays in shared memory (decoding a task is a binary search: dependent loads, once per visit)
    constexpr int PLAN_S = 32;
    __shared__ int s_plan[4 * PLAN_S + 2 + 9 * PLAN_S + 1];
    __shared__ float s_temp[PLAN_S];
    if (plan.n_sweeps <= PLAN_S) {
        const int S = plan.n_sweeps;
        int *sox = s_plan, *soy = s_plan + PLAN_S + 1, *slo = s_plan + 2 * PLAN_S + 2, *shi = s_plan + 3 * PLAN_S + 2, *sbase = s_plan + 4 * PLAN_S + 2;
        for (int k = threadIdx.x; k <= S; k += 32 * NW) { sox[k] = plan.ox[k]; soy[k] = plan.oy[k]; }
        for (int k = threadIdx.x; k < S; k += 32 * NW) { slo[k] = plan.wi_lo[k]; shi[k] = plan.wi_hi[k]; s_temp[k] = plan.temp[k]; }
        for (int k = threadIdx.x; k <= 9 * S; k += 32 * NW) sbase[k] = plan.task_base[k];
        plan.ox = sox; plan.oy = soy; plan.wi_lo = slo; plan.wi_hi = shi; plan.task_base = sbase; plan.temp = s_temp;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_task = atomicAdd(plan.next_task, 1);
    for (;;) {
        __syncthreads();  // s_task is set; the previous visit's statistics have been read out of `w` / `sc`
        const int t = s_task;
        if (t >= plan.total_tasks) break;
        // decode (sweep, colour, scene, window)
        int lo = 0, hi = 9 * plan.n_sweeps - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (plan.task_base[mid] <= t) lo = mid; else hi = mid - 1; }
        const int s = lo / 9, col = lo % 9, ci = col / 3, cj = col % 3;
        const int ox = plan.ox[s + 1], oy = plan.oy[s + 1];
        const int wlo = plan.wi_lo[s], whi = plan.wi_hi[s];
        const int first_i = wlo + (ci - wlo % 3 + 3) % 3;
        // copy the scene descriptor (needs H, W for the window counts; all scenes of a batch have the same shape)
        const int H = scenes[0].c.H, W = scenes[0].c.W;
        const int nwx = (H + ox + 31) / 32, nwy = (W + oy + 31) / 32;
        const int a_j = cj < nwy ? (nwy - cj + 2) / 3 : 0;
        const int a_i = first_i < whi ? (whi - first_i + 2) / 3 : 0;
        const int per_scene = a_i * a_j, l = t - plan.task_base[lo];
        const int scene = l / per_scene, local = l - scene * per_scene;
        const int wi = first_i + 3 * (local / a_j), wj = cj + 3 * (local % a_j);
        {
            const int *src = reinterpret_cast<const int *>(scenes + scene);
            int *dst = reinterpret_cast<int *>(&sc);
            for (int k = threadIdx.x; k < (int)(sizeof(SceneDev<R>) / 4); k += 32 * NW) dst[k] = src[k];
        }
        __syncthreads();
        const int stamp = plan.stamp0 + s;
        const size_t gsz = (size_t)plan.dg * plan.dg;
        // wait for the conflicting earlier visits (possibly completed by a neighbour GPU)
        if (warp == 0) {
            const int *done_cur = sc.done + (size_t)(stamp & 1) * gsz, *done_prev = sc.done + (size_t)((stamp + 1) & 1) * gsz;
            const int px0 = 32 * wi - ox, py0 = 32 * wj - oy;
            const int x0 = max(px0, 0), x1 = min(px0 + 32, H), y0 = max(py0, 0), y1 = min(py0 + 32, W);
            int pi0 = 0, pi1 = -1, pj0 = 0, pj1 = -1;
            if (stamp > 1) {
                const int oxp = plan.ox[s], oyp = plan.oy[s];
                const int nwxp = (H + oxp + 31) / 32, nwyp = (W + oyp + 31) / 32;
                pi0 = max((max(x0 - 64, 0) + oxp) >> 5, 0); pi1 = min((min(x1 - 1 + 64, H - 1) + oxp) >> 5, nwxp - 1);
                pj0 = max((max(y0 - 64, 0) + oyp) >> 5, 0); pj1 = min((min(y1 - 1 + 64, W - 1) + oyp) >> 5, nwyp - 1);
            }
            const int pw = pj1 - pj0 + 1, pn = (pi1 - pi0 + 1) * pw;
            const long long t_wait = clock64();
            for (;;) {
                bool ok = true;
                if (lane < 25) {  // same sweep, earlier colours, |dwi|, |dwj| <= 2
                    const int ni = wi + lane / 5 - 2, nj = wj + lane % 5 - 2;
                    if (ni >= 0 && nj >= 0 && ni < nwx && nj < nwy && (ni % 3) * 3 + nj % 3 < col) {
                        const int *p = done_cur + ni * plan.dg + nj;
                        ok = (SPLIT ? ld_relaxed_sys(p) : ld_relaxed(p)) >= stamp;
                    }
                }
                for (int q = lane; q < pn; q += 32) {  // every window of the previous sweep within 64 px
                    const int *p = done_prev + (pi0 + q / pw) * plan.dg + pj0 + q % pw;
                    ok = ok && (SPLIT ? ld_relaxed_sys(p) : ld_relaxed(p)) >= stamp - 1;
                }
                if (__all_sync(MPP_FULL, ok)) {
                    // polls are relaxed loads (see ld_relaxed: an acquire per poll empties the L1 of the SM each time); across
                    // GPUs one acquire at the end of the wait keeps the formal ordering with the neighbour's release
                    if (SPLIT && lane == 0) (void)ld_acquire_sys(done_cur + wi * plan.dg + wj);
                    break;
                }
                __nanosleep(SPLIT ? 200 : 100);
                // watchdog: a dependency that never completes (a neighbour rank that did not launch, bands of one device that are
                // not co-resident) must end as an error, not as a hung GPU: ~4 s at 2 GHz (30 s across ranks, which start at different times)
                if (clock64() - t_wait > (SPLIT ? 60000000000LL : 8000000000LL)) { if (lane == 0) atomicOr(sc.c.err, ERRF_TIMEOUT); break; }
            }
        }
        // a split scene runs alone on its rank: its context is also in the kernel's parameter bank (`scene0`), where the inlined
        // code reads it as immediate constant operands (see k_windows_dataflow); a batch of tiles has one context per scene
        if (SPLIT) window_visit<R, NW, DBG, false, SPLIT>(scene0.c, sc.c, w, scratch, wi, wj, ox, oy, per_visit, plan.temp[s], sc.seed, sweep_offset + (uint64_t)s, 0u, dbg_maxdiff);
        else window_visit<R, NW, DBG, false, SPLIT>(sc.c, sc.c, w, scratch, wi, wj, ox, oy, per_visit, plan.temp[s], sc.seed, sweep_offset + (uint64_t)s, 0u, dbg_maxdiff);
        if (threadIdx.x == 32 % (32 * NW)) s_task = atomicAdd(plan.next_task, 1);  // (see k_windows_dataflow)
        __syncthreads();
        if (threadIdx.x == 0) {
            const size_t idx = (size_t)(stamp & 1) * gsz + (size_t)wi * plan.dg + wj;
            if (SPLIT) {
                __threadfence_system();
                const int x0 = max(32 * wi - ox, 0), x1 = min(32 * wi - ox + 32, H);
                st_release_sys(sc.done + idx, stamp);
                if (sc.done_up && x0 < sc.notify_lo) st_release_sys(sc.done_up + idx, stamp);
                if (sc.done_down && x1 > sc.notify_hi) st_release_sys(sc.done_down + idx, stamp);
            } else {
                st_release(sc.done + idx, stamp);  // (one MEMBAR: see k_windows_dataflow)
            }
        }
        visit_statistics<R, NW, false>(sc.c, w, per_visit);
    }
}
