// B200-native Marked-Point-Process RJMCMC hot path: kernels + the C ABI of include/mpp_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "mpp_device.cuh"
#include "mpp_proposals.cuh"
#include "mpp_chain.cuh"
#include "mpp_sweep2.cuh"
#include "mpp_multi.cuh"
#include "mpp_split_merge.cuh"

// ================================================================================================ host ctx
struct mpp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int precision = 0;
    int H = 0, W = 0, nx = 0, ny = 0, ncell = 0;
    uint32_t *d_mask = nullptr;
    void *d_recs = nullptr;
    double *d_cell_cdf = nullptr;
    double *d_rowcum = nullptr;         // [H][W+1] row prefix sums of det
    float *d_marksum = nullptr;         // [3][H][W] class sums of the mark rows
    int *d_scan = nullptr;              // [ncell + 1] exclusive prefix of per-cell populations
    int *d_nobj = nullptr;
    int *d_rowcount = nullptr;          // [nx] objects per row of cells (global uniform pick)
    uint32_t *d_next_uid = nullptr;
    uint32_t *d_err = nullptr;
    unsigned long long *d_counters = nullptr;
    unsigned long long *d_kstats = nullptr;  // [MPP_WINDOW_STATS]
    unsigned char *d_nms_state = nullptr;  // [H*W] naive-init scratch, allocated on first use
    const float *det = nullptr;
    const float *marks = nullptr;
    int map_row0 = 0, map_rows = 0;     // rows of the scene the maps cover (mpp_set_maps: all; mpp_set_maps_band: a band)
    float det_sum = 0.f;
    bool maps_set = false, model_set = false, kernels_set = false;
    ModelDev m;
    KernDev k;
    float visit_alpha = 1.f, visit_tfloor = 0.f;  // temperature decay inside a window visit (set per mpp_run_windows call)
    int *h_pinned = nullptr;            // small pinned read-back area (16 x 8 bytes)
    void *d_plan = nullptr;             // device scratch of the dataflow schedule (offsets, temperatures, completion grids)
    size_t plan_bytes = 0;
    int num_sms = 0;
    uint32_t window_uid_next = 0x80000000u;  // uids of objects born in mpp_run_windows: host-tracked, upper half of the uid space
    // multi-scene / split-scene schedule (mpp_run_windows_batch)
    int *d_done = nullptr;                   // [2][dg][dg] completion stamps (own cudaMalloc: exported to the neighbour ranks)
    int sweeps_done = 0, last_ox = 0, last_oy = 0;  // sweeps stamped so far, grid offset of the last one
    bool split = false;                      // this context holds one band of a scene (mpp_split_attach*)
    int row_lo = 0, row_hi = 0;              // ... the pixel rows of the band
    uint32_t *mask_up = nullptr, *mask_down = nullptr;
    void *recs_up = nullptr, *recs_down = nullptr;
    int *done_up = nullptr, *done_down = nullptr;
    void *ipc_opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // peer mappings to close on detach
    mpp_window_trace *trace = nullptr;       // per-proposal trace of the window sampler (mpp_set_window_trace)
    unsigned long long trace_capacity = 0, trace_sweep0 = 0;
};

#define MPP_MAX_DEVICES 64
#ifdef MPP_TRACE
#define MPP_KSTATS_COPY 64   // instrumented build: entries 40.. hold per-kernel evaluation timers (tools/visit_timers.py)
#else
#define MPP_KSTATS_COPY MPP_WINDOW_STATS
#endif
#define MPP_KSTATS_ALLOC 64
static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(MPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));              \
    } while (0)

template <typename R>
static Ctx<R> device_view(const mpp_ctx *h) {
    Ctx<R> c;
    c.H = h->H; c.W = h->W; c.nx = h->nx; c.ny = h->ny; c.ncell = h->ncell;
    c.mask = h->d_mask;
    c.recs = reinterpret_cast<Rec<R> *>(h->d_recs);
    // band-local maps: bias the pointers so that the kernels keep indexing by scene rows
    c.det = h->det - (ptrdiff_t)h->map_row0 * h->W;
    c.marks = h->marks - (ptrdiff_t)h->map_row0 * h->W * MPP_N_CLASSES;
    c.mark_plane = (size_t)(h->map_rows ? h->map_rows : h->H) * h->W * MPP_N_CLASSES;
    c.det_sum = h->det_sum;
    c.cell_cdf = h->d_cell_cdf;
    c.rowcum = h->d_rowcum;
    c.marksum = h->d_marksum;
    c.n_objects = h->d_nobj; c.next_uid = h->d_next_uid; c.err = h->d_err; c.counters = h->d_counters; c.kstats = h->d_kstats;
    c.m = h->m; c.k = h->k;
    c.visit_alpha = h->visit_alpha; c.visit_tfloor = h->visit_tfloor;
    c.trace = h->trace; c.trace_capacity = h->trace_capacity; c.trace_sweep0 = h->trace_sweep0;
    c.own_lo = h->split ? h->row_lo / MPP_CELL_SIZE : 0;
    c.own_hi = h->split ? (h->row_hi + MPP_CELL_SIZE - 1) / MPP_CELL_SIZE : h->nx;
    c.mask_up = h->mask_up; c.mask_down = h->mask_down;
    c.recs_up = reinterpret_cast<Rec<R> *>(h->recs_up); c.recs_down = reinterpret_cast<Rec<R> *>(h->recs_down);
    return c;
}

#define DISPATCH(h, CALL)                                         \
    do {                                                          \
        if ((h)->precision == MPP_PRECISION_FP64) { typedef double R; CALL; } \
        else { typedef float R; CALL; }                           \
    } while (0)

static int check_device_errors(mpp_ctx *h) {
    CUDA_TRY(cudaMemcpyAsync(h->h_pinned, h->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const uint32_t f = *reinterpret_cast<uint32_t *>(h->h_pinned);
    if (!f) return MPP_OK;
    CUDA_TRY(cudaMemsetAsync(h->d_err, 0, sizeof(uint32_t), h->stream));
    if (f & ERRF_OUT_OF_BOUNDS) return fail(MPP_ERR_OUT_OF_BOUNDS, "object outside the support (point_set.py:99)");
    if (f & ERRF_NOT_FOUND) return fail(MPP_ERR_NOT_FOUND, "removal of an object that is not in the set");
    if (f & ERRF_CELL_FULL) return fail(MPP_ERR_CELL_FULL, "more than 32 objects in one 32x32 cell");
    if (f & ERRF_TIMEOUT) return fail(MPP_ERR_TIMEOUT, "a window visit waited too long for a visit it depends on (neighbour rank not running?)");
    return fail(MPP_ERR_NEIGHBOURHOOD, "more than MPP_KMAX objects around one perturbation");
}

// ================================================================================================ kernels
// ---- K5a: per-cell detection mass (warp per cell) -------------------------------------------------
__global__ void k_cell_mass(const float *__restrict__ det, int H, int W, int ny, int ncell, double *__restrict__ mass) {
    const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (cell >= ncell) return;
    const int x = (cell / ny) * 32 + lane, y0 = (cell % ny) * 32;
    double s = 0.0;
    if (x < H) {
        const float *row = det + (size_t)x * W + y0;
        const int wy = min(32, W - y0);
        for (int j = 0; j < wy; ++j) s += (double)__ldg(row + j);
    }
    s = warp_sum(s);
    if (lane == 0) mass[cell] = s;
}

// K5b: exclusive prefix sums of every row of det (warp per row)
__global__ void k_row_prefix(const float *__restrict__ det, int H, int W, double *__restrict__ rowcum) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= H) return;
    const float *src = det + (size_t)row * W;
    double *dst = rowcum + (size_t)row * ((size_t)W + 1);
    double base = 0.0;
    if (lane == 0) dst[0] = 0.0;
    for (int b = 0; b < W; b += 32) {
        const double v = (b + lane < W) ? (double)__ldg(src + b + lane) : 0.0;
        const double incl = warp_incl_scan(v, lane);
        if (b + lane < W) dst[b + lane + 1] = base + incl;
        base += __shfl_sync(MPP_FULL, incl, 31);
    }
}

// K5c: class sums of the mark rows: one warp per 32 consecutive (mark, pixel) rows, coalesced 128-byte row reads
__global__ void k_mark_sums(const float *__restrict__ marks, size_t n_rows, float *__restrict__ out) {
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const size_t r0 = warp * 32;
    float mine = 0.f;
    for (int k = 0; k < 32; ++k) {
        const size_t r = r0 + k;
        if (r >= n_rows) break;
        const float s = warp_sum(__ldg(marks + r * MPP_N_CLASSES + lane));
        if (lane == k) mine = s;
    }
    if (r0 + lane < n_rows) out[r0 + lane] = mine;
}

// single-block inclusive scan of doubles in place (ncell <= a few 1e5)
__global__ void k_inv_total(double *cdf, int n) { cdf[n] = cdf[n - 1] > 0.0 ? 1.0 / cdf[n - 1] : 0.0; }

__global__ void k_scan_double(double *v, int n) {
    __shared__ double part[1024];
    const int t = threadIdx.x, per = (n + blockDim.x - 1) / blockDim.x;
    const int b = t * per, e = min(n, b + per);
    double s = 0.0;
    for (int i = b; i < e; ++i) s += v[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) { double a = 0.0; for (int i = 0; i < (int)blockDim.x; ++i) { double x = part[i]; part[i] = a; a += x; } }
    __syncthreads();
    double a = part[t];
    for (int i = b; i < e; ++i) { a += v[i]; v[i] = a; }
}

// ---- K1: objects ----------------------------------------------------------------------------------
template <typename R>
__global__ void k_add_objects(Ctx<R> c, const int32_t *__restrict__ xy, const double *__restrict__ marks,
                              const uint32_t *__restrict__ cls, const uint32_t *__restrict__ uid, int n,
                              uint32_t *__restrict__ out_handle) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = xy[2 * i], y = xy[2 * i + 1];
    if (out_handle) out_handle[i] = MPP_NO_OBJECT;
    if (x < 0 || y < 0 || x >= c.H || y >= c.W) { atomicOr(c.err, ERRF_OUT_OF_BOUNDS); return; }
    const R size = (R)marks[3 * i], ratio = (R)marks[3 * i + 1], angle = (R)marks[3 * i + 2];
    const uint32_t pc = cls ? cls[i] : pack_cls(value_to_class<R>(0, size), value_to_class<R>(1, ratio), value_to_class<R>(2, angle));
    const uint32_t id = uid ? uid[i] : atomicAdd(c.next_uid, 1u);
    const Rec<R> r = make_rec(c, x, y, size, ratio, angle, pc, id);
    const int cell = cell_of(c, x, y);
    for (;;) {
        const uint32_t msk = atomicOr(c.mask + cell, 0u);
        if (msk == 0xffffffffu) { atomicOr(c.err, ERRF_CELL_FULL); return; }
        const int slot = __ffs(~msk) - 1;
        const uint32_t old = atomicOr(c.mask + cell, 1u << slot);
        if (!(old & (1u << slot))) {
            store_rec(c.recs + (size_t)cell * 32 + slot, r);
            if (out_handle) out_handle[i] = (uint32_t)cell * 32u + slot;
            atomicAdd(c.n_objects, 1);
            return;
        }
    }
}

template <typename R>
__global__ void k_remove_objects(Ctx<R> c, const uint32_t *__restrict__ handle, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t h = handle[i];
    if ((h >> 5) >= (uint32_t)c.ncell) { atomicOr(c.err, ERRF_NOT_FOUND); return; }
    const uint32_t bit = 1u << (h & 31);
    const uint32_t old = atomicAnd(c.mask + (h >> 5), ~bit);
    if (old & bit) atomicSub(c.n_objects, 1);
    else atomicOr(c.err, ERRF_NOT_FOUND);
}

// single-block exclusive scan of the per-cell populations -> scan[0..ncell]
__global__ void k_scan_population(const uint32_t *__restrict__ mask, int ncell, int *__restrict__ scan) {
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (ncell + blockDim.x - 1) / blockDim.x;
    const int b = min(ncell, t * per), e = min(ncell, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += __popc(mask[i]);
    part[t] = s;
    __syncthreads();
    if (t == 0) { int a = 0; for (int i = 0; i < (int)blockDim.x; ++i) { int x = part[i]; part[i] = a; a += x; } scan[ncell] = a; }
    __syncthreads();
    int a = part[t];
    for (int i = b; i < e; ++i) { scan[i] = a; a += __popc(mask[i]); }
}

template <typename R>
__global__ void k_read_objects(Ctx<R> c, const int *__restrict__ scan, int capacity, uint32_t *handle, int32_t *xy,
                               double *marks, uint32_t *uid) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= c.ncell) return;
    uint32_t msk = c.mask[cell];
    int pos = scan[cell];
    while (msk) {
        const int slot = __ffs(msk) - 1;
        msk &= msk - 1;
        if (pos < capacity) {
            const Rec<R> r = load_rec(c.recs + (size_t)cell * 32 + slot);
            if (handle) handle[pos] = (uint32_t)cell * 32u + slot;
            if (xy) { xy[2 * pos] = r.x; xy[2 * pos + 1] = r.y; }
            if (marks) { marks[3 * pos] = (double)r.size; marks[3 * pos + 1] = (double)r.ratio; marks[3 * pos + 2] = (double)r.angle; }
            if (uid) uid[pos] = r.uid;
        }
        ++pos;
    }
}

// ---- K2-K4: per-object energy vectors (warp per object) ---------------------------------------------
template <typename R>
__global__ void k_energy_vectors(Ctx<R> c, const uint32_t *__restrict__ handle, int n, double *__restrict__ out_vectors,
                                 double *__restrict__ out_combined) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Scratch<R> &s = reinterpret_cast<Scratch<R> *>(smem)[wib];
    const int i = blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= n) return;
    const uint32_t h = handle[i];
    const bool ok = (h >> 5) < (uint32_t)c.ncell && ((c.mask[h >> 5] >> (h & 31)) & 1u);
    if (!ok) {
        if (lane == 0) {
            atomicOr(c.err, ERRF_NOT_FOUND);
            if (out_combined) out_combined[i] = 0.0;
        }
        return;
    }
    const Rec<R> u = load_rec(c.recs + h);
    const Terms<R> t = warp_object_terms(c, s, h, u, lane);
    if (lane == 0) {
        R v[MPP_MAX_TERMS];
        term_vector(c.m, t, v);
        if (out_vectors)
            for (int k = 0; k < MPP_MAX_TERMS; ++k) out_vectors[(size_t)i * MPP_MAX_TERMS + k] = k < c.m.n_terms ? (double)v[k] : 0.0;
        if (out_combined) out_combined[i] = (double)combine(c.m, t);
    }
}

// deterministic totals: {sum of all vector entries, sum of combined}
__global__ void k_totals(const double *__restrict__ vectors, const double *__restrict__ combined, int n, double *__restrict__ out) {
    __shared__ double pa[256], pb[256];
    const int t = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int i = t; i < n; i += blockDim.x) {
        if (vectors) for (int k = 0; k < MPP_MAX_TERMS; ++k) a += vectors[(size_t)i * MPP_MAX_TERMS + k];
        if (combined) b += combined[i];
    }
    pa[t] = a; pb[t] = b;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (t < o) { pa[t] += pa[t + o]; pb[t] += pb[t + o]; }
        __syncthreads();
    }
    if (t == 0) { out[0] = pa[0]; out[1] = pb[0]; }
}

// ---- K3: batch of independent perturbations (warp per perturbation) ---------------------------------
template <typename R>
__device__ __forceinline__ bool resolve_proposal(const Ctx<R> &c, const mpp_proposal &p, int lane, bool *has_rem,
                                                 uint32_t *rem_handle, Rec<R> *rem, bool *has_add, Rec<R> *add) {
    *has_rem = p.rem_uid != MPP_NO_OBJECT;
    *has_add = p.add_uid != MPP_NO_OBJECT;
    *rem_handle = MPP_NO_OBJECT;
    if (*has_rem) {
        *rem_handle = find_by_uid(c, p.rem_x, p.rem_y, p.rem_uid, lane);
        if (*rem_handle == MPP_NO_OBJECT) { if (lane == 0) atomicOr(c.err, ERRF_NOT_FOUND); return false; }
        *rem = load_rec(c.recs + *rem_handle);
    }
    if (*has_add) {
        if (p.add_x < 0 || p.add_y < 0 || p.add_x >= c.H || p.add_y >= c.W) { if (lane == 0) atomicOr(c.err, ERRF_OUT_OF_BOUNDS); return false; }
        *add = make_rec(c, p.add_x, p.add_y, (R)p.add_size, (R)p.add_ratio, (R)p.add_angle, p.add_cls, p.add_uid);
    }
    return true;
}

template <typename R>
__global__ void k_delta_batch(Ctx<R> c, const mpp_proposal *__restrict__ props, int m, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Scratch<R> &s = reinterpret_cast<Scratch<R> *>(smem)[wib];
    const int i = blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= m) return;
    const mpp_proposal p = props[i];
    bool has_rem, has_add;
    uint32_t rh;
    Rec<R> rem, add;
    if (!resolve_proposal(c, p, lane, &has_rem, &rh, &rem, &has_add, &add)) { if (lane == 0) out[i] = 0.0; return; }
    const R d = warp_delta(c, s, has_rem, rh, rem, has_add, add, lane);
    if (lane == 0) out[i] = (double)d;
}

// ---- K7: sequential replay of a recorded proposal stream (one warp) --------------------------------
template <typename R>
__global__ void k_replay(Ctx<R> c, const mpp_proposal *__restrict__ props, int m, double t0, double alpha_t,
                         double t_target, mpp_step_result *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    Scratch<R> &s = *reinterpret_cast<Scratch<R> *>(smem);
    const int lane = threadIdx.x & 31;
    int n = __ldcg(c.n_objects);
    double temp = t0;
    for (int step = 0; step < m; ++step) {
        const mpp_proposal p = props[step];
        bool has_rem, has_add;
        uint32_t rh;
        Rec<R> rem, add;
        mpp_step_result res;
        res.delta_e = 0; res.fwd = 0; res.bwd = 0; res.log_alpha = 0; res.temperature = temp; res.accepted = -1; res.n_after = n;
        if (resolve_proposal(c, p, lane, &has_rem, &rh, &rem, &has_add, &add)) {
            const R d = warp_delta(c, s, has_rem, rh, rem, has_add, add, lane);
            double fwd, bwd;
            proposal_probs(c, p.kernel, has_rem, rem, has_add, add, p.delta0, p.delta1, p.param_id, p.new_class, (double)n,
                           c.k.intensity, lane, &fwd, &bwd);
            const double la = (-(double)d / temp) + log(bwd + MPP_EPS) - log(fwd + MPP_EPS);  // rjmcmc.py:105-107
            const bool acc = log(p.u + MPP_EPS) < la;                                         // rjmcmc.py:113
            if (acc) {  // EPointsSet.apply_perturbation: removal first (energy_point_set.py:123-141)
                if (has_rem) { erase_handle(c, rh, lane); --n; }
                __syncwarp();
                if (has_add) { insert_rec(c, add, lane); ++n; }
                __syncwarp();
            }
            res.delta_e = (double)d; res.fwd = fwd; res.bwd = bwd; res.log_alpha = la; res.accepted = acc ? 1 : 0; res.n_after = n;
        }
        if (lane == 0) out[step] = res;
        if (temp > t_target) temp *= alpha_t;  // rjmcmc.py:158-159
    }
    if (lane == 0) *c.n_objects = n;
}

// ---- sequential device chain with the reference's global kernels (R15-R21) ---------------------------
__global__ void k_row_counts(const uint32_t *__restrict__ mask, int nx, int ny, int *__restrict__ row_count, int *__restrict__ n_objects) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nx) return;
    int s = 0;
    for (int j = lane; j < ny; j += 32) s += __popc(mask[row * ny + j]);
    s = warp_sum(s);
    if (lane == 0) { row_count[row] = s; if (n_objects) atomicAdd(n_objects, s); }
}

template <typename R>
__device__ __forceinline__ void fill_proposal(mpp_proposal *p, int kernel, const Drawn<R> &d, double u) {
    p->kernel = kernel;
    p->rem_x = d.has_rem ? d.rem.x : 0; p->rem_y = d.has_rem ? d.rem.y : 0;
    p->rem_uid = d.has_rem ? d.rem.uid : MPP_NO_OBJECT;
    p->add_x = d.has_add ? d.add.x : 0; p->add_y = d.has_add ? d.add.y : 0;
    p->add_uid = d.has_add ? 0u : MPP_NO_OBJECT;
    p->add_cls = d.has_add ? d.add.cls : 0u;
    p->add_size = d.has_add ? (double)d.add.size : 0.0;
    p->add_ratio = d.has_add ? (double)d.add.ratio : 0.0;
    p->add_angle = d.has_add ? (double)d.add.angle : 0.0;
    p->delta0 = d.d0; p->delta1 = d.d1; p->param_id = d.param_id; p->new_class = d.new_class; p->u = u;
}

template <typename R>
__global__ void k_sample_proposals(Ctx<R> c, const int *__restrict__ row_count, const int32_t *__restrict__ kernel_ids, int m,
                                   uint64_t seed, uint64_t offset, mpp_proposal *__restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= m) return;
    const uint64_t id = offset + (uint64_t)i;
    Philox rng(seed, (uint32_t)id, (uint32_t)(id >> 32), 0x5a3du);
    const uint4 r0 = rng.next(), r1 = rng.next(), r2 = rng.next();
    const int n = __ldcg(c.n_objects);
    int kernel = kernel_ids ? kernel_ids[i] : -1;
    if (kernel < 0 || kernel > 7) kernel = draw_kernel(c.k, u01(r0.x, r0.y));
    Drawn<R> d;
    draw_global(c, row_count, n, kernel, r0, r1, r2, lane, &d);
    if (lane == 0) fill_proposal(out + i, kernel, d, u01(r2.z, r2.w));
}

template <typename R>
__global__ void k_proposal_probs(Ctx<R> c, const mpp_proposal *__restrict__ props, int m, double *__restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= m) return;
    const mpp_proposal p = props[i];
    bool has_rem, has_add;
    uint32_t rh;
    Rec<R> rem, add;
    double fwd = 0, bwd = 0;
    if (resolve_proposal(c, p, lane, &has_rem, &rh, &rem, &has_add, &add))
        proposal_probs(c, p.kernel, has_rem, rem, has_add, add, p.delta0, p.delta1, p.param_id, p.new_class,
                       (double)__ldcg(c.n_objects), c.k.intensity, lane, &fwd, &bwd);
    if (lane == 0) { out[2 * i] = fwd; out[2 * i + 1] = bwd; }
}

template <typename R>
__global__ void k_run_chain(Ctx<R> c, int *__restrict__ row_count, int n_steps, double t0, double alpha_t, double t_target,
                            uint64_t seed, uint64_t step_offset, mpp_step_result *__restrict__ trace) {
    extern __shared__ __align__(16) unsigned char smem[];
    Scratch<R> &s = *reinterpret_cast<Scratch<R> *>(smem);
    const int lane = threadIdx.x & 31;
    int n = __ldcg(c.n_objects);
    double temp = t0;
    unsigned n_acc = 0, n_birth = 0, n_death = 0;
    for (int step = 0; step < n_steps; ++step) {
        const uint64_t id = step_offset + (uint64_t)step;
        Philox rng(seed, (uint32_t)id, (uint32_t)(id >> 32), 0xc4a1u);
        const uint4 r0 = rng.next(), r1 = rng.next(), r2 = rng.next();
        const int kernel = draw_kernel(c.k, u01(r0.x, r0.y));
        Drawn<R> d;
        draw_global(c, row_count, n, kernel, r0, r1, r2, lane, &d);
        bool valid = true;
        if (d.has_add) {  // capacity of the destination cell (MPP_CELL_CAPACITY slots)
            uint32_t dm = __ldcg(c.mask + cell_of(c, d.add.x, d.add.y));
            if (d.has_rem && (d.rem_handle >> 5) == (uint32_t)cell_of(c, d.add.x, d.add.y)) dm &= ~(1u << (d.rem_handle & 31));
            if (dm == 0xffffffffu) { valid = false; if (lane == 0) atomicOr(c.err, ERRF_CELL_FULL); }
        }
        double de = 0, fwd = 0, bwd = 0, la = 0;
        bool acc = false;
        if (valid) {
            de = (double)warp_delta(c, s, d.has_rem, d.rem_handle, d.rem, d.has_add, d.add, lane);
            proposal_probs(c, kernel, d.has_rem, d.rem, d.has_add, d.add, d.d0, d.d1, d.param_id, d.new_class, (double)n,
                           c.k.intensity, lane, &fwd, &bwd);
            la = (-de / temp) + log(bwd + MPP_EPS) - log(fwd + MPP_EPS);      // rjmcmc.py:105-107
            acc = log(u01(r2.z, r2.w) + MPP_EPS) < la;                        // rjmcmc.py:113
            if (acc) {
                if (d.has_rem) {
                    erase_handle(c, d.rem_handle, lane);
                    if (lane == 0) row_count[d.rem.x >> 5] -= 1;
                    --n;
                }
                __syncwarp();
                if (d.has_add) {
                    if (lane == 0) d.add.uid = atomicAdd(c.next_uid, 1u);
                    d.add.uid = __shfl_sync(MPP_FULL, d.add.uid, 0);
                    insert_rec(c, d.add, lane);
                    if (lane == 0) row_count[d.add.x >> 5] += 1;
                    ++n;
                }
                __threadfence();
                __syncwarp();
                ++n_acc;
                if (d.has_add && !d.has_rem) ++n_birth;
                if (d.has_rem && !d.has_add) ++n_death;
            }
        }
        if (trace && lane == 0) {
            mpp_step_result res;
            res.delta_e = de; res.fwd = fwd; res.bwd = bwd; res.log_alpha = la; res.temperature = temp;
            res.accepted = acc ? 1 : 0; res.n_after = n;
            trace[step] = res;
        }
        if (temp > t_target) temp *= alpha_t;  // rjmcmc.py:158-159
    }
    if (lane == 0) {
        *c.n_objects = n;
        atomicAdd(c.counters + 0, (unsigned long long)n_steps);
        atomicAdd(c.counters + 1, (unsigned long long)n_acc);
        atomicAdd(c.counters + 2, (unsigned long long)n_birth);
        atomicAdd(c.counters + 3, (unsigned long long)n_death);
        atomicAdd(c.counters + 4, (unsigned long long)n_steps);
    }
}

// ---- R14: combinator over given energy vectors ------------------------------------------------------
__global__ void k_combine(ModelDev m, const double *__restrict__ vectors, int n, double *__restrict__ out_per, double *__restrict__ out_total) {
    __shared__ double part[256];
    const int t = threadIdx.x;
    double acc = 0.0;
    for (int i = t; i < n; i += blockDim.x) {
        double v[MPP_MAX_TERMS];
        for (int k = 0; k < MPP_MAX_TERMS; ++k) v[k] = vectors[(size_t)i * MPP_MAX_TERMS + k];
        const double e = combine_v<double>(m, v);
        if (out_per) out_per[i] = e;
        acc += e;
    }
    part[t] = acc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (t < o) part[t] += part[t + o]; __syncthreads(); }
    if (t == 0 && out_total) *out_total = part[0];
}

// ---- K5 test entry: draw births from the data-driven sampler ---------------------------------------
template <typename R>
__global__ void k_sample_births(Ctx<R> c, int n, uint64_t seed, int32_t *__restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    Philox rng(seed, (uint32_t)i, 0x5a17u, 0);
    const uint4 a = rng.next(), b = rng.next();
    int x, y;
    sample_birth_pixel(c, u01(a.x, a.y), u01f(a.z), u01f(a.w), lane, &x, &y);
    const int c0 = sample_mark_class(c, 0, x, y, u01f(b.x), lane);
    const int c1 = sample_mark_class(c, 1, x, y, u01f(b.y), lane);
    const int c2 = sample_mark_class(c, 2, x, y, u01f(b.z), lane);
    if (lane == 0) { out[5 * i] = x; out[5 * i + 1] = y; out[5 * i + 2] = c0; out[5 * i + 3] = c1; out[5 * i + 4] = c2; }
}

// ---- K6: parallel sweep over one colour class (warp per active cell) -------------------------------
template <typename R>
__global__ void k_sweep(Ctx<R> c, int stride, int ci, int cj, int n_ai, int n_aj, int per_visit, double temp,
                        uint64_t seed, uint64_t sweep_id) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Scratch<R> &s = reinterpret_cast<Scratch<R> *>(smem)[wib];
    const int a = blockIdx.x * (blockDim.x >> 5) + wib;
    if (a >= n_ai * n_aj) return;
    const int i = ci + stride * (a / n_aj), j = cj + stride * (a % n_aj);
    const int cell = j + i * c.ny;
    const int x0 = i * 32, x1 = min(x0 + 32, c.H), y0 = j * 32, y1 = min(y0 + 32, c.W);
    const int wx = x1 - x0, wy = y1 - y0;
    const bool in_cell_moves = stride < 4;  // stride 3: moves must stay in the cell; stride >= 4: |move| <= 16 px
    const double q_unif = ((double)wx * (double)wy) / ((double)c.H * (double)c.W);
    const double cell_mass = c.cell_cdf[cell] - (cell > 0 ? c.cell_cdf[cell - 1] : 0.0);
    const double q_data = cell_mass / c.cell_cdf[c.ncell - 1];
    Philox rng(seed, (uint32_t)cell, (uint32_t)sweep_id, (uint32_t)(sweep_id >> 32));
    unsigned n_acc = 0, n_birth = 0, n_death = 0, n_eval = 0;
    int dn = 0;
    double cdf[8];
    { double acc = 0; for (int k = 0; k < 8; ++k) { acc += c.k.p[k]; cdf[k] = acc; } }

    for (int it = 0; it < per_visit; ++it) {
        const uint4 r0 = rng.next(), r1 = rng.next(), r2 = rng.next();
        const uint32_t msk = __ldcg(c.mask + cell);
        const int nc = __popc(msk);
        int kernel = 7;
        { const double uk = u01(r0.x, r0.y) * cdf[7]; for (int k = 6; k >= 0; --k) if (uk < cdf[k]) kernel = k; }
        bool has_rem = false, has_add = false, valid = true;
        uint32_t rh = MPP_NO_OBJECT;
        Rec<R> rem, add;
        double d0 = 0, d1 = 0, lambda = 1.0;
        int param_id = 0, new_class = 0;
        if (kernel >= 1 && kernel != 2 && nc > 0) {  // kernels that pick an existing object uniformly in the cell
            const int pick = min(nc - 1, (int)(u01f(r0.z) * (float)nc));
            rh = (uint32_t)cell * 32u + __fns(msk, 0, pick + 1);
            rem = load_rec(c.recs + rh);
            has_rem = true;
        }
        switch (kernel) {
        case 0: {  // UniformRectangleSampler.sample shape_samplers.py:136-141, restricted to the cell
            const int x = x0 + min(wx - 1, (int)(u01f(r0.z) * (float)wx)), y = y0 + min(wy - 1, (int)(u01f(r0.w) * (float)wy));
            const R size = (R)(u01(r1.x, r1.y) * 32.0), ratio = (R)u01(r1.z, r1.w), angle = (R)(u01(r2.x, r2.y) * 3.14159265358979323846);
            add = make_rec(c, x, y, size, ratio, angle,
                           pack_cls(value_to_class<R>(0, size), value_to_class<R>(1, ratio), value_to_class<R>(2, angle)), 0u);
            has_add = true;
            lambda = c.k.intensity * q_unif;
            break;
        }
        case 2: {  // RectangleSampler.sample shape_samplers.py:90-98, restricted to the cell
            if (!(cell_mass > 0.0)) { valid = false; break; }
            int x, y;
            sample_window(c, x0, x1, y0, y1, u01f(r0.z), u01f(r0.w), lane, &x, &y, nullptr);
            const int c0 = sample_mark_class(c, 0, x, y, u01f(r1.x), lane);
            const int c1 = sample_mark_class(c, 1, x, y, u01f(r1.y), lane);
            const int c2 = sample_mark_class(c, 2, x, y, u01f(r1.z), lane);
            add = make_rec(c, x, y, mark_edge<R>(0, c0), mark_edge<R>(1, c1), mark_edge<R>(2, c2), pack_cls(c0, c1, c2), 0u);
            has_add = true;
            lambda = c.k.intensity * q_data;
            break;
        }
        case 1: lambda = c.k.intensity * q_unif; break;
        case 3: lambda = c.k.intensity * q_data; if (has_rem && !(cell_mass > 0.0)) valid = false; break;
        case 4: if (has_rem) {  // GaussianTranslationKernel.sample_perturbation transform_kernels.py:24-36
            box_muller(r1.x, r1.y, r1.z, r1.w, &d0, &d1);
            d0 *= c.k.trl_sigma; d1 *= c.k.trl_sigma;
            int nx_ = (int)((double)rem.x + d0), ny_ = (int)((double)rem.y + d1);  // astype(int): truncation
            nx_ = min(max(nx_, 0), c.H - 1); ny_ = min(max(ny_, 0), c.W - 1);
            add = rem; add.x = nx_; add.y = ny_;
            has_add = true;
        } break;
        case 5: if (has_rem) {  // DataDrivenTranslationKernel.sample_perturbation :77-89
            const int md = c.k.trl_max_delta;
            int x, y;
            sample_window(c, max(0, rem.x - md), min(rem.x + md + 1, c.H), max(0, rem.y - md), min(rem.y + md + 1, c.W),
                          u01f(r1.x), u01f(r1.y), lane, &x, &y, nullptr);
            add = rem; add.x = x; add.y = y;
            has_add = true;
        } break;
        case 6: if (has_rem) {  // GaussianShapeTransformKernel.sample_perturbation :128-143
            param_id = min(2, (int)(u01f(r0.w) * 3.0f));
            double dummy;
            box_muller(r1.x, r1.y, r1.z, r1.w, &d0, &dummy);
            d0 *= c.k.trf_sigma[param_id];
            R v = (param_id == 0 ? rem.size : (param_id == 1 ? rem.ratio : rem.angle)) + (R)d0;
            const R vmax = (R)mark_vmax(param_id);
            if (param_id == 2) v = v - r_floor(v / vmax) * vmax;  // python % on a cyclic mark
            else v = r_min(r_max(v, (R)0), vmax);
            if (param_id == 2 && !(v < vmax)) v = 0;
            add = rem;
            if (param_id == 0) add.size = v; else if (param_id == 1) add.ratio = v; else add.angle = v;
            const int nc_ = value_to_class<R>(param_id, v);
            add.cls = (rem.cls & ~(0xffu << (8 * param_id))) | ((uint32_t)nc_ << (8 * param_id));
            fill_geometry(add);
            has_add = true;
        } break;
        default: if (has_rem) {  // DataDrivenShapeTransformKernel.sample_perturbation :179-200
            param_id = min(2, (int)(u01f(r0.w) * 3.0f));
            new_class = sample_mark_class(c, param_id, rem.x, rem.y, u01f(r1.x), lane);
            const R v = mark_edge<R>(param_id, new_class);
            add = rem;
            if (param_id == 0) add.size = v; else if (param_id == 1) add.ratio = v; else add.angle = v;
            add.cls = (rem.cls & ~(0xffu << (8 * param_id))) | ((uint32_t)new_class << (8 * param_id));
            fill_geometry(add);
            has_add = true;
        } break;
        }
        if (has_add && has_rem) {
            if (kernel == 4 || kernel == 5) {
                const bool stays = add.x >= x0 && add.x < x1 && add.y >= y0 && add.y < y1;
                const bool short_move = abs(add.x - rem.x) <= 16 && abs(add.y - rem.y) <= 16;
                if (in_cell_moves ? !stays : !short_move) valid = false;  // symmetric restriction -> plain rejection
                if (valid) fill_unit_energies(c, add);
            } else {
                fill_unit_energies(c, add);  // classes changed
            }
        }
        if (valid && has_add) {  // a full destination cell rejects the proposal (capacity MPP_CELL_CAPACITY)
            const int dcell = cell_of(c, add.x, add.y);
            uint32_t dm = __ldcg(c.mask + dcell);
            if (has_rem && (rh >> 5) == (uint32_t)dcell) dm &= ~(1u << (rh & 31));
            if (dm == 0xffffffffu) valid = false;
        }
        if (valid && (has_rem || has_add)) {
            ++n_eval;
            const R d = warp_delta(c, s, has_rem, rh, rem, has_add, add, lane);
            double fwd, bwd;
            proposal_probs(c, kernel, has_rem, rem, has_add, add, d0, d1, param_id, new_class, (double)nc, lambda, lane, &fwd, &bwd);
            const double la = (-(double)d / temp) + log(bwd + MPP_EPS) - log(fwd + MPP_EPS);
            const bool acc = log(u01(r2.z, r2.w) + MPP_EPS) < la;
            if (acc) {
                if (has_rem) erase_handle(c, rh, lane);
                __syncwarp();
                if (has_add) {
                    if (lane == 0) add.uid = atomicAdd(c.next_uid, 1u);
                    add.uid = __shfl_sync(MPP_FULL, add.uid, 0);
                    insert_rec(c, add, lane);
                }
                __syncwarp();
                ++n_acc;
                if (has_add && !has_rem) { ++n_birth; ++dn; }
                if (has_rem && !has_add) { ++n_death; --dn; }
            }
        }
    }
    if (lane == 0) {
        atomicAdd(c.counters + 0, (unsigned long long)per_visit);
        atomicAdd(c.counters + 1, (unsigned long long)n_acc);
        atomicAdd(c.counters + 2, (unsigned long long)n_birth);
        atomicAdd(c.counters + 3, (unsigned long long)n_death);
        atomicAdd(c.counters + 4, (unsigned long long)n_eval);
        if (dn) atomicAdd(c.n_objects, dn);
    }
}

// ---- R4: neighbour query (PointsSet.get_potential_neighbors / get_neighbors, point_set.py:111-149) ----------
template <typename R>
__global__ void k_query_neighbors(Ctx<R> c, int x, int y, int max_offset, long long r2, int euclidean, uint32_t exclude,
                                  int capacity, uint32_t *__restrict__ out, int *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int iu = x >> 5, ju = y >> 5;
    const int i0 = max(iu - max_offset, 0), i1 = min(iu + max_offset, c.nx - 1);
    const int j0 = max(ju - max_offset, 0), j1 = min(ju + max_offset, c.ny - 1);
    int total = 0;
    for (int i = i0; i <= i1; ++i)
        for (int j = j0; j <= j1; ++j) {
            const int cell = j + i * c.ny;
            const uint32_t msk = c.mask[cell];
            bool in = (msk >> lane) & 1u;
            const uint32_t h = (uint32_t)cell * 32u + lane;
            if (in && h == exclude) in = false;
            if (in && euclidean) {
                const Rec<R> *p = c.recs + h;
                const long long dx = p->x - x, dy = p->y - y;
                in = dx * dx + dy * dy <= r2;
            }
            const uint32_t b = __ballot_sync(MPP_FULL, in);
            if (in) { const int pos = total + __popc(b & ((1u << lane) - 1)); if (pos < capacity) out[pos] = h; }
            total += __popc(b);
        }
    if (lane == 0) *count = total;
}

template <typename R>
__global__ void k_pair_values(Ctx<R> c, const uint32_t *__restrict__ ha, const uint32_t *__restrict__ hb, int n,
                              double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Scratch<R> &s = reinterpret_cast<Scratch<R> *>(smem)[wib];
    const int i = blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= n) return;
    const uint32_t a = ha[i], b = hb[i];
    const bool ok = (a >> 5) < (uint32_t)c.ncell && (b >> 5) < (uint32_t)c.ncell && ((c.mask[a >> 5] >> (a & 31)) & 1u) &&
                    ((c.mask[b >> 5] >> (b & 31)) & 1u);
    if (!ok) { if (lane == 0) { atomicOr(c.err, ERRF_NOT_FOUND); out[2 * i] = nan(""); out[2 * i + 1] = nan(""); } return; }
    const Rec<R> ra = load_rec(c.recs + a), rb = load_rec(c.recs + b);
    const int dx = ra.x - rb.x, dy = ra.y - rb.y, d2 = dx * dx + dy * dy;
    const Geo<R> ga = geo_of(ra), gb = geo_of(rb);
    const R ov = d2 <= c.m.ov_d2 ? pair_overlap(c.m, ga, gb, d2, s.clipx + lane, s.clipy + lane) : (R)0;
    if (lane == 0) {
        out[2 * i] = d2 <= c.m.ov_d2 ? (double)ov : nan("");
        const R al = (c.m.rewarding ? (R)-1 : (R)1) * align_magnitude(ga, gb, c.m.rewarding);
        out[2 * i + 1] = d2 <= c.m.al_d2 ? (double)al : nan("");
    }
}

// ---- naive initial configuration (sample_rjmcmc.py:23-35) -------------------------------------------
// Greedy distance NMS (utils/nms.py:68-109) == maximal independent set by priority: a pixel is kept iff no
// kept pixel of higher priority lies within `thr`.  Solved in rounds: an undecided candidate that beats every
// undecided candidate around it is kept, then everything around a kept pixel is suppressed.
// state: 0 not a candidate, 1 undecided, 2 kept, 3 suppressed.  Priority = (score, flat index).
__global__ void k_nms_threshold(const float *__restrict__ det, int n, float thr, unsigned char *__restrict__ state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) state[i] = det[i] >= thr ? 1 : 0;
}
__global__ void k_nms_select(const float *__restrict__ det, int H, int W, int rad, int rad2, unsigned char *state, int *changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W || state[i] != 1) return;
    const int x = i / W, y = i % W;
    const float sc = det[i];
    for (int dx = -rad; dx <= rad; ++dx)
        for (int dy = -rad; dy <= rad; ++dy) {
            if ((dx | dy) == 0 || dx * dx + dy * dy > rad2) continue;
            const int px = x + dx, py = y + dy;
            if (px < 0 || py < 0 || px >= H || py >= W) continue;
            const int j = px * W + py;
            const unsigned char st = state[j];
            if (st == 2) return;  // will be suppressed by k_nms_suppress
            if (st == 1) { const float o = det[j]; if (o > sc || (o == sc && j > i)) return; }
        }
    state[i] = 2;  // no undecided neighbour can ever beat it
    *changed = 1;
}
__global__ void k_nms_suppress(int H, int W, int rad, int rad2, unsigned char *state, int *remaining) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W || state[i] != 1) return;
    const int x = i / W, y = i % W;
    for (int dx = -rad; dx <= rad; ++dx)
        for (int dy = -rad; dy <= rad; ++dy) {
            if (dx * dx + dy * dy > rad2) continue;
            const int px = x + dx, py = y + dy;
            if (px < 0 || py < 0 || px >= H || py >= W) continue;
            if (state[px * W + py] == 2) { state[i] = 3; return; }
        }
    *remaining = 1;
}
template <typename R>
__global__ void k_nms_insert(Ctx<R> c, const unsigned char *__restrict__ state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.H * c.W || state[i] != 2) return;
    const int x = i / c.W, y = i % c.W;
    int cl[3];
    for (int k = 0; k < 3; ++k) {  // np.argmax over the raw mark rows (first maximum)
        const float *row = mark_row(c, k, x, y);
        int best = 0; float bv = row[0];
        for (int b = 1; b < MPP_N_CLASSES; ++b) { const float v = row[b]; if (v > bv) { bv = v; best = b; } }
        cl[k] = best;
    }
    const Rec<R> r = make_rec(c, x, y, mark_edge<R>(0, cl[0]), mark_edge<R>(1, cl[1]), mark_edge<R>(2, cl[2]),
                              pack_cls(cl[0], cl[1], cl[2]), atomicAdd(c.next_uid, 1u));
    const int cell = cell_of(c, x, y);
    for (;;) {
        const uint32_t msk = atomicOr(c.mask + cell, 0u);
        if (msk == 0xffffffffu) { atomicOr(c.err, ERRF_CELL_FULL); return; }
        const int slot = __ffs(~msk) - 1;
        const uint32_t old = atomicOr(c.mask + cell, 1u << slot);
        if (!(old & (1u << slot))) { store_rec(c.recs + (size_t)cell * 32 + slot, r); atomicAdd(c.n_objects, 1); return; }
    }
}

// ---- multi-GPU halo: pack / unpack the objects of a band of rows -----------------------------------
template <typename R>
__global__ void k_pack_rows(Ctx<R> c, int row_lo, int row_hi, double *__restrict__ buf, int capacity, int *__restrict__ count) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= c.ncell) return;
    const int cx0 = (cell / c.ny) * 32;
    if (cx0 + 32 <= row_lo || cx0 >= row_hi) return;
    uint32_t msk = c.mask[cell];
    while (msk) {
        const int slot = __ffs(msk) - 1;
        msk &= msk - 1;
        const Rec<R> r = load_rec(c.recs + (size_t)cell * 32 + slot);
        if (r.x < row_lo || r.x >= row_hi) continue;
        const int pos = atomicAdd(count, 1);
        if (pos < capacity) {
            double *o = buf + (size_t)pos * 8;
            o[0] = r.x; o[1] = r.y; o[2] = (double)r.size; o[3] = (double)r.ratio; o[4] = (double)r.angle;
            o[5] = (double)r.cls; o[6] = (double)r.uid; o[7] = 0.0;
        }
    }
}
template <typename R>
__global__ void k_erase_rows(Ctx<R> c, int row_lo, int row_hi) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= c.ncell) return;
    const int cx0 = (cell / c.ny) * 32;
    if (cx0 + 32 <= row_lo || cx0 >= row_hi) return;
    uint32_t msk = c.mask[cell], keep = msk;
    int removed = 0;
    while (msk) {
        const int slot = __ffs(msk) - 1;
        msk &= msk - 1;
        const int x = c.recs[(size_t)cell * 32 + slot].x;
        if (x >= row_lo && x < row_hi) { keep &= ~(1u << slot); ++removed; }
    }
    if (removed) { c.mask[cell] = keep; atomicSub(c.n_objects, removed); }
}
template <typename R>
__global__ void k_unpack_rows(Ctx<R> c, const double *__restrict__ buf, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *o = buf + (size_t)i * 8;
    const int x = (int)o[0], y = (int)o[1];
    if (x < 0 || y < 0 || x >= c.H || y >= c.W) { atomicOr(c.err, ERRF_OUT_OF_BOUNDS); return; }
    const Rec<R> r = make_rec(c, x, y, (R)o[2], (R)o[3], (R)o[4], (uint32_t)o[5], (uint32_t)o[6]);
    const int cell = cell_of(c, x, y);
    for (;;) {
        const uint32_t msk = atomicOr(c.mask + cell, 0u);
        if (msk == 0xffffffffu) { atomicOr(c.err, ERRF_CELL_FULL); return; }
        const int slot = __ffs(~msk) - 1;
        const uint32_t old = atomicOr(c.mask + cell, 1u << slot);
        if (!(old & (1u << slot))) { store_rec(c.recs + (size_t)cell * 32 + slot, r); atomicAdd(c.n_objects, 1); return; }
    }
}

// ================================================================================================ C ABI
static const int WARPS_PER_BLOCK = 4;

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

extern "C" {

int mpp_abi_version(void) { return MPP_ABI_VERSION; }
const char *mpp_last_error(void) { return g_last_error.c_str(); }
int mpp_abi_struct_size(int which) {
    switch (which) {
    case 0: return (int)sizeof(mpp_model_params);
    case 1: return (int)sizeof(mpp_kernel_params);
    case 2: return (int)sizeof(mpp_proposal);
    case 3: return (int)sizeof(mpp_step_result);
    case 4: return (int)sizeof(mpp_window_trace);
    case 5: return (int)sizeof(mpp_split_merge);
    default: return -1;
    }
}

int mpp_ctx_create(mpp_ctx **out, int device, int height, int width, int precision, void *stream) {
    if (!out || height <= 0 || width <= 0 || device < 0 || device >= MPP_MAX_DEVICES) return fail(MPP_ERR_INVALID, "mpp_ctx_create: bad arguments");
    if (precision != MPP_PRECISION_FP32 && precision != MPP_PRECISION_FP64) return fail(MPP_ERR_INVALID, "mpp_ctx_create: precision");
    CUDA_TRY(cudaSetDevice(device));
    mpp_ctx *h = new (std::nothrow) mpp_ctx();
    if (!h) return fail(MPP_ERR_INVALID, "out of host memory");
    h->device = device; h->stream = (cudaStream_t)stream; h->precision = precision;
    h->H = height; h->W = width;
    h->nx = (height + MPP_CELL_SIZE - 1) / MPP_CELL_SIZE;  // point_set.py:60-61
    h->ny = (width + MPP_CELL_SIZE - 1) / MPP_CELL_SIZE;
    h->ncell = h->nx * h->ny;
    const size_t rec = precision == MPP_PRECISION_FP64 ? sizeof(Rec<double>) : sizeof(Rec<float>);
    CUDA_TRY(cudaMalloc(&h->d_mask, sizeof(uint32_t) * h->ncell));
    CUDA_TRY(cudaMalloc(&h->d_recs, rec * (size_t)h->ncell * MPP_CELL_CAPACITY));
    CUDA_TRY(cudaMalloc(&h->d_cell_cdf, sizeof(double) * (h->ncell + 1)));  // [ncell]: 1 / total mass
    CUDA_TRY(cudaMalloc(&h->d_rowcum, sizeof(double) * (size_t)height * ((size_t)width + 1)));
    CUDA_TRY(cudaMalloc(&h->d_marksum, sizeof(float) * 3 * (size_t)height * (size_t)width));
    CUDA_TRY(cudaMalloc(&h->d_scan, sizeof(int) * (h->ncell + 1)));
    CUDA_TRY(cudaMalloc(&h->d_nobj, sizeof(int)));
    CUDA_TRY(cudaMalloc(&h->d_rowcount, sizeof(int) * h->nx));
    CUDA_TRY(cudaMalloc(&h->d_next_uid, sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&h->d_err, sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&h->d_counters, sizeof(unsigned long long) * 8));
    CUDA_TRY(cudaMalloc(&h->d_kstats, sizeof(unsigned long long) * MPP_KSTATS_ALLOC));
    CUDA_TRY(cudaMemsetAsync(h->d_kstats, 0, sizeof(unsigned long long) * MPP_KSTATS_ALLOC, h->stream));
    CUDA_TRY(cudaMallocHost(&h->h_pinned, 128));
    CUDA_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaMemsetAsync(h->d_mask, 0, sizeof(uint32_t) * h->ncell, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_recs, 0, rec * (size_t)h->ncell * MPP_CELL_CAPACITY, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_nobj, 0, sizeof(int), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_next_uid, 0, sizeof(uint32_t), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_err, 0, sizeof(uint32_t), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 8, h->stream));
    memset(&h->m, 0, sizeof(h->m));
    memset(&h->k, 0, sizeof(h->k));
    // dynamic shared memory opt-in for the warp-scratch kernels
    CUDA_TRY(set_smem(k_energy_vectors<float>, WARPS_PER_BLOCK * sizeof(Scratch<float>)));
    CUDA_TRY(set_smem(k_energy_vectors<double>, WARPS_PER_BLOCK * sizeof(Scratch<double>)));
    CUDA_TRY(set_smem(k_delta_batch<float>, WARPS_PER_BLOCK * sizeof(Scratch<float>)));
    CUDA_TRY(set_smem(k_delta_batch<double>, WARPS_PER_BLOCK * sizeof(Scratch<double>)));
    CUDA_TRY(set_smem(k_sweep<float>, WARPS_PER_BLOCK * sizeof(Scratch<float>)));
    CUDA_TRY(set_smem(k_sweep<double>, WARPS_PER_BLOCK * sizeof(Scratch<double>)));
    CUDA_TRY(set_smem(k_pair_values<float>, WARPS_PER_BLOCK * sizeof(Scratch<float>)));
    CUDA_TRY(set_smem(k_pair_values<double>, WARPS_PER_BLOCK * sizeof(Scratch<double>)));
    CUDA_TRY(set_smem(k_replay<float>, sizeof(Scratch<float>)));
    CUDA_TRY(set_smem(k_replay<double>, sizeof(Scratch<double>)));
    *out = h;
    return MPP_OK;
}

int mpp_ctx_destroy(mpp_ctx *h) {
    if (!h) return MPP_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_mask); cudaFree(h->d_recs); cudaFree(h->d_cell_cdf); cudaFree(h->d_rowcum); cudaFree(h->d_marksum); cudaFree(h->d_scan); cudaFree(h->d_nobj);
    for (void *p : h->ipc_opened) if (p) cudaIpcCloseMemHandle(p);
    cudaFree(h->d_done);
    cudaFree(h->d_rowcount); cudaFree(h->d_plan); cudaFree(h->d_next_uid); cudaFree(h->d_err); cudaFree(h->d_counters); cudaFree(h->d_kstats); cudaFree(h->d_nms_state);
    cudaFreeHost(h->h_pinned);
    delete h;
    return MPP_OK;
}

int mpp_ctx_reset(mpp_ctx *h, void *stream) {
    if (!h) return fail(MPP_ERR_INVALID, "mpp_ctx_reset: null ctx");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(h->d_mask, 0, sizeof(uint32_t) * h->ncell, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_nobj, 0, sizeof(int), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_next_uid, 0, sizeof(uint32_t), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_err, 0, sizeof(uint32_t), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 8, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_kstats, 0, sizeof(unsigned long long) * MPP_KSTATS_ALLOC, h->stream));
    h->det = nullptr; h->marks = nullptr; h->det_sum = 0.f; h->map_row0 = 0; h->map_rows = 0;
    h->maps_set = false; h->model_set = false; h->kernels_set = false;
    h->window_uid_next = 0x80000000u;
    h->trace = nullptr; h->trace_capacity = 0; h->trace_sweep0 = 0;
    h->sweeps_done = 0; h->last_ox = 0; h->last_oy = 0;
    for (void *&p : h->ipc_opened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    h->split = false; h->row_lo = 0; h->row_hi = 0;
    h->mask_up = h->mask_down = nullptr; h->recs_up = h->recs_down = nullptr; h->done_up = h->done_down = nullptr;
    if (h->d_done) { const int dg = (std::max(h->H, h->W) + 31) / 32 + 2; CUDA_TRY(cudaMemsetAsync(h->d_done, 0, sizeof(int) * 2 * dg * dg, h->stream)); }
    memset(&h->m, 0, sizeof(h->m));
    memset(&h->k, 0, sizeof(h->k));
    return MPP_OK;
}

int mpp_set_maps(mpp_ctx *h, const float *det, const float *marks, double det_sum) {
    if (!h || !det || !marks) return fail(MPP_ERR_INVALID, "mpp_set_maps: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    h->det = det; h->marks = marks; h->map_row0 = 0; h->map_rows = h->H;
    const int blocks = (h->ncell + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    k_cell_mass<<<blocks, WARPS_PER_BLOCK * 32, 0, h->stream>>>(det, h->H, h->W, h->ny, h->ncell, h->d_cell_cdf);
    k_scan_double<<<1, 1024, 0, h->stream>>>(h->d_cell_cdf, h->ncell);
    k_inv_total<<<1, 1, 0, h->stream>>>(h->d_cell_cdf, h->ncell);
    k_row_prefix<<<(h->H + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, h->stream>>>(det, h->H, h->W, h->d_rowcum);
    {
        const size_t n_rows = (size_t)3 * h->H * h->W, n_warps = (n_rows + 31) / 32;
        k_mark_sums<<<(unsigned)((n_warps + 7) / 8), 256, 0, h->stream>>>(marks, n_rows, h->d_marksum);
    }
    CUDA_TRY(cudaGetLastError());
    if (det_sum > 0.0) {
        h->det_sum = (float)det_sum;
    } else {
        double *tmp = reinterpret_cast<double *>(h->h_pinned);
        CUDA_TRY(cudaMemcpyAsync(tmp, h->d_cell_cdf + (h->ncell - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->det_sum = (float)*tmp;
    }
    h->maps_set = true;
    return MPP_OK;
}

// ---- utils/sampler2d.py:5-48: weighted pixel draws from a density map (stand-alone: no context needed) ------
__global__ void k_row_totals(const double *__restrict__ rowcum, int H, int W, double *__restrict__ tot) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < H) tot[r] = rowcum[(size_t)r * ((size_t)W + 1) + W];
}
__global__ void k_sample_points_2d(const double *__restrict__ rowcum, const double *__restrict__ row_cdf, int H, int W, int n, uint64_t seed,
                                   int32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox rng(seed, (uint32_t)i, (uint32_t)((uint64_t)i >> 32), 0x5a2d0000u);
    const uint4 q = rng.next();
    // inverse CDF: first row whose inclusive prefix exceeds u * total, then first column of that row likewise
    // (numpy Generator.choice: cdf = cumsum(p); searchsorted(u * cdf[-1], side='right'))
    const double t = u01(q.x, q.y) * row_cdf[H - 1];
    int lo = 0, hi = H - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (row_cdf[mid] > t) hi = mid; else lo = mid + 1; }
    const double *rc = rowcum + (size_t)lo * ((size_t)W + 1);
    const double tc = u01(q.z, q.w) * rc[W];
    int a = 0, b = W - 1;
    while (a < b) { const int mid = (a + b) >> 1; if (rc[mid + 1] > tc) b = mid; else a = mid + 1; }
    out[2 * i] = lo; out[2 * i + 1] = a;
}

int mpp_sample_points_2d(const float *density, int height, int width, int n, uint64_t seed, int32_t *out_xy, double *scratch, int device,
                         void *stream) {
    if (!density || !out_xy || !scratch || height < 1 || width < 1 || n < 0) return fail(MPP_ERR_INVALID, "mpp_sample_points_2d: bad arguments");
    if (n == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    double *rowcum = scratch, *row_cdf = scratch + (size_t)height * ((size_t)width + 1);
    k_row_prefix<<<(height + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, st>>>(density, height, width, rowcum);
    k_row_totals<<<(height + 255) / 256, 256, 0, st>>>(rowcum, height, width, row_cdf);
    k_scan_double<<<1, 1024, 0, st>>>(row_cdf, height);
    k_sample_points_2d<<<(n + 127) / 128, 128, 0, st>>>(rowcum, row_cdf, height, width, n, seed, out_xy);
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

__global__ void k_set_inv_total(double *cdf, int n, double total) { cdf[n] = total > 0.0 ? 1.0 / total : 0.0; }

int mpp_set_maps_band(mpp_ctx *h, const float *det_band, const float *marks_band, int row0, int rows, double det_sum_scene) {
    if (!h || !det_band || !marks_band) return fail(MPP_ERR_INVALID, "mpp_set_maps_band: null argument");
    if (row0 < 0 || rows < 1 || row0 + rows > h->H || !(det_sum_scene > 0.0)) return fail(MPP_ERR_INVALID, "mpp_set_maps_band: bad band or scene sum");
    CUDA_TRY(cudaSetDevice(h->device));
    h->det = det_band; h->marks = marks_band; h->map_row0 = row0; h->map_rows = rows;
    k_set_inv_total<<<1, 1, 0, h->stream>>>(h->d_cell_cdf, h->ncell, det_sum_scene);
    k_row_prefix<<<(rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, h->stream>>>(det_band, rows, h->W,
                                                                                                       h->d_rowcum + (size_t)row0 * ((size_t)h->W + 1));
    for (int i = 0; i < 3; ++i) {
        const size_t n_rows = (size_t)rows * h->W, n_warps = (n_rows + 31) / 32;
        k_mark_sums<<<(unsigned)((n_warps + 7) / 8), 256, 0, h->stream>>>(marks_band + (size_t)i * rows * h->W * MPP_N_CLASSES, n_rows,
                                                                        h->d_marksum + (size_t)i * h->H * h->W + (size_t)row0 * h->W);
    }
    CUDA_TRY(cudaGetLastError());
    h->det_sum = (float)det_sum_scene;
    h->maps_set = true;
    return MPP_OK;
}

}  // extern "C" (helpers)

static int model_from_params(const mpp_model_params *p, ModelDev &m) {
    if (!p) return fail(MPP_ERR_INVALID, "mpp_set_model: null argument");
    if (p->setup != MPP_SETUP_LEGACY && p->setup != MPP_SETUP_NO_CALIBRATION && p->setup != MPP_SETUP_TOY)
        return fail(MPP_ERR_INVALID, "mpp_set_model: setup");
    if (p->setup == MPP_SETUP_TOY && (p->combinator != MPP_COMB_RAW_SUM || p->toy_pair_value < 0.0))
        return fail(MPP_ERR_INVALID, "mpp_set_model: the toy setup takes the raw-sum combinator and a non-negative pair value");
    if (p->combinator < 0 || p->combinator > MPP_COMB_MANUAL_HIERARCHICAL) return fail(MPP_ERR_INVALID, "mpp_set_model: combinator");
    if (p->combinator == MPP_COMB_HIERARCHICAL && p->setup != MPP_SETUP_LEGACY)
        return fail(MPP_ERR_INVALID, "hierarchical combinator needs the legacy term names (hierarchical.py:22-29)");
    if (p->overlap_max_dist > MPP_CELL_SIZE || p->align_max_dist > MPP_CELL_SIZE)
        return fail(MPP_ERR_INVALID, "interaction distances above the 32-px cell size are not supported");
    m.setup = p->setup; m.comb = p->combinator; m.ratio_prior = p->ratio_prior; m.rewarding = p->rewarding;
    m.n_terms = p->setup == MPP_SETUP_LEGACY ? 5 : (p->setup == MPP_SETUP_TOY ? 2 : (p->ratio_prior ? 8 : 7));
    m.ov_d2 = p->overlap_max_dist < 0 ? -1 : (int)floor(p->overlap_max_dist * p->overlap_max_dist);
    m.al_d2 = (p->setup == MPP_SETUP_TOY || p->align_max_dist < 0) ? -1 : (int)floor(p->align_max_dist * p->align_max_dist);
    m.premapped = p->marks_are_energies ? 1 : 0;
    m.toy_unit = p->toy_unit_value; m.toy_pair = p->toy_pair_value;
    {
        const double t2 = p->toy_pair_dist * p->toy_pair_dist;
        m.toy_d2 = p->toy_pair_dist < 0 ? -1 : (p->toy_pair_strict ? (int)ceil(t2) - 1 : (int)floor(t2));
    }
    m.max_d2 = m.ov_d2 > m.al_d2 ? m.ov_d2 : m.al_d2;
    m.pos_thr = (float)p->pos_threshold;
    for (int i = 0; i < 3; ++i) { m.coef[i] = (float)p->remap_coef[i]; m.icpt[i] = (float)p->remap_intercept[i]; }
    m.min_area = p->min_area; m.max_area = p->max_area; m.target_ratio = p->target_ratio;
    for (int i = 0; i < MPP_MAX_TERMS; ++i) m.w[i] = p->comb_w[i];
    m.bias = p->comb_bias; m.thr = p->comb_threshold;
    {   // fold the combinator into the gated linear form used by the sweep kernels (see ModelDev)
        const int n = m.n_terms;
        double a0 = 0, b[MPP_MAX_TERMS] = {0, 0, 0, 0, 0, 0, 0, 0}, c0 = 0;
        m.gate = 0; m.logistic = 0; m.gate_thr = (float)p->comb_threshold;
        switch (p->combinator) {
        case MPP_COMB_HIERARCHICAL:
            a0 = m.w[5] * m.w[0]; b[1] = m.w[5] * m.w[1]; b[2] = m.w[6] * m.w[2]; b[3] = m.w[6] * m.w[3]; b[4] = m.w[6] * m.w[4];
            c0 = m.bias; m.gate = 1;
            break;
        case MPP_COMB_MANUAL_HIERARCHICAL:
            a0 = m.w[0]; for (int k = 1; k < n; ++k) b[k] = m.w[k];
            m.gate = 1;
            break;
        case MPP_COMB_LOGISTIC:
            a0 = m.w[0]; for (int k = 1; k < n; ++k) b[k] = m.w[k];
            c0 = (double)n * m.bias; m.logistic = 1;
            break;
        default:
            a0 = 1.0; for (int k = 1; k < n; ++k) b[k] = 1.0;
            break;
        }
        m.c_pos = (float)a0; m.c_0 = (float)c0;
        m.c_m0 = m.c_m1 = m.c_m2 = m.c_ov = m.c_al = m.c_area = m.c_ratio = 0.f;
        if (p->setup == MPP_SETUP_LEGACY) { m.c_m0 = (float)b[1]; m.c_ov = (float)b[2]; m.c_al = (float)b[3]; m.c_area = (float)b[4]; }
        else if (p->setup == MPP_SETUP_TOY) { m.c_ov = (float)b[1]; }
        else {
            m.c_m0 = (float)b[1]; m.c_m1 = (float)b[2]; m.c_m2 = (float)b[3]; m.c_ov = (float)b[4]; m.c_al = (float)b[5];
            m.c_area = (float)b[6]; m.c_ratio = p->ratio_prior ? (float)b[7] : 0.f;
        }
        m.f_min_area = (float)p->min_area; m.f_max_area = (float)p->max_area; m.f_target_ratio = (float)p->target_ratio;
    }
    return MPP_OK;
}

int mpp_set_model(mpp_ctx *h, const mpp_model_params *p) {
    if (!h) return fail(MPP_ERR_INVALID, "mpp_set_model: null ctx");
    ModelDev m;
    memset(&m, 0, sizeof(m));
    const int rc = model_from_params(p, m);
    if (rc != MPP_OK) return rc;
    h->m = m;
    h->model_set = true;
    return MPP_OK;
}

extern "C" {

int mpp_set_kernels(mpp_ctx *h, const mpp_kernel_params *p) {
    if (!h || !p) return fail(MPP_ERR_INVALID, "mpp_set_kernels: null argument");
    if (!(p->intensity > 0.0)) return fail(MPP_ERR_INVALID, "mpp_set_kernels: intensity must be > 0");
    if (p->data_translation_max_delta < 0 || p->data_translation_max_delta > 15) return fail(MPP_ERR_INVALID, "mpp_set_kernels: max_delta in [0,15]");
    KernDev &k = h->k;
    for (int i = 0; i < 8; ++i) { k.p[i] = p->p_kernel[i]; k.pf[i] = (float)p->p_kernel[i]; }
    k.pk_e0 = (float)(k.p[0] / (k.p[0] + k.p[2])); k.pk_e2 = (float)(k.p[2] / (k.p[0] + k.p[2]));
    k.unif_scale = p->intensity / ((double)h->H * (double)h->W);
    k.intensity = p->intensity;
    k.trl_sigma = p->gauss_translation_sigma;
    k.trl_max_delta = p->data_translation_max_delta;
    const double range[3] = {32.0, 1.0, 3.14159265358979323846};  // transform_kernels.py:122
    for (int i = 0; i < 3; ++i) k.trf_sigma[i] = p->gauss_transform_sigma * range[i];
    h->kernels_set = true;
    return MPP_OK;
}

static int refresh_row_counts_impl(mpp_ctx *h);
static int refresh_row_counts(mpp_ctx *h) { return refresh_row_counts_impl(h); }
#define NEED(h, cond, what) do { if (!(h)) return fail(MPP_ERR_INVALID, "null ctx"); if (!(cond)) return fail(MPP_ERR_STATE, what); } while (0)

int mpp_add_objects(mpp_ctx *h, const int32_t *xy, const double *marks, const uint32_t *cls, const uint32_t *uid, int n,
                    uint32_t *out_handle) {
    NEED(h, h->model_set && (h->maps_set || h->m.setup == MPP_SETUP_TOY), "mpp_add_objects: set maps and model first");
    if (n < 0 || (n > 0 && (!xy || !marks))) return fail(MPP_ERR_INVALID, "mpp_add_objects: bad arguments");
    if (n == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    DISPATCH(h, (k_add_objects<R><<<(n + 127) / 128, 128, 0, h->stream>>>(device_view<R>(h), xy, marks, cls, uid, n, out_handle)));
    CUDA_TRY(cudaGetLastError());
    return check_device_errors(h);
}

int mpp_remove_objects(mpp_ctx *h, const uint32_t *handle, int n) {
    NEED(h, true, "");
    if (n < 0 || (n > 0 && !handle)) return fail(MPP_ERR_INVALID, "mpp_remove_objects: bad arguments");
    if (n == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    DISPATCH(h, (k_remove_objects<R><<<(n + 127) / 128, 128, 0, h->stream>>>(device_view<R>(h), handle, n)));
    CUDA_TRY(cudaGetLastError());
    return check_device_errors(h);
}

int mpp_clear_objects(mpp_ctx *h) {
    NEED(h, true, "");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemsetAsync(h->d_mask, 0, sizeof(uint32_t) * h->ncell, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_nobj, 0, sizeof(int), h->stream));
    return MPP_OK;
}

int mpp_num_objects(mpp_ctx *h, int *n_host) {
    NEED(h, n_host != nullptr, "mpp_num_objects: null output");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->split) {  // a neighbour rank's windows add to / remove from this band's boundary cells: recount from the masks
        const int rc = refresh_row_counts(h);
        if (rc != MPP_OK) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(h->h_pinned, h->d_nobj, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_host = *h->h_pinned;
    return MPP_OK;
}

int mpp_read_objects(mpp_ctx *h, int capacity, uint32_t *handle, int32_t *xy, double *marks, uint32_t *uid, int *n_host) {
    NEED(h, true, "");
    if (capacity < 0) return fail(MPP_ERR_INVALID, "mpp_read_objects: capacity");
    CUDA_TRY(cudaSetDevice(h->device));
    k_scan_population<<<1, 1024, 0, h->stream>>>(h->d_mask, h->ncell, h->d_scan);
    if (capacity > 0)
        DISPATCH(h, (k_read_objects<R><<<(h->ncell + 127) / 128, 128, 0, h->stream>>>(device_view<R>(h), h->d_scan, capacity, handle, xy, marks, uid)));
    CUDA_TRY(cudaGetLastError());
    if (n_host) {
        CUDA_TRY(cudaMemcpyAsync(h->h_pinned, h->d_scan + h->ncell, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        *n_host = *h->h_pinned;
    }
    return MPP_OK;
}

int mpp_query_neighbors(mpp_ctx *h, int x, int y, double radius, int euclidean, uint32_t exclude_handle, int capacity,
                        uint32_t *out_handle, int *n_host) {
    NEED(h, n_host != nullptr, "mpp_query_neighbors: null output");
    if (capacity < 0 || (capacity > 0 && !out_handle) || radius < 0) return fail(MPP_ERR_INVALID, "mpp_query_neighbors: bad arguments");
    if (x < 0 || y < 0 || x >= h->H || y >= h->W) return fail(MPP_ERR_OUT_OF_BOUNDS, "object outside the support (point_set.py:99)");
    CUDA_TRY(cudaSetDevice(h->device));
    const int max_offset = (int)ceil(radius / (double)MPP_CELL_SIZE);  // point_set.py:129
    const long long r2 = (long long)floor(radius * radius);
    int *count = h->d_scan;
    DISPATCH(h, (k_query_neighbors<R><<<1, 32, 0, h->stream>>>(device_view<R>(h), x, y, max_offset, r2, euclidean, exclude_handle,
                                                             capacity, out_handle, count)));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h->h_pinned, count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_host = *h->h_pinned;
    return MPP_OK;
}

int mpp_copy_state(mpp_ctx *dst, const mpp_ctx *src) {
    if (!dst || !src) return fail(MPP_ERR_INVALID, "mpp_copy_state: null ctx");
    if (dst->H != src->H || dst->W != src->W || dst->precision != src->precision || dst->device != src->device)
        return fail(MPP_ERR_INVALID, "mpp_copy_state: contexts differ in shape, precision or device");
    CUDA_TRY(cudaSetDevice(dst->device));
    const size_t rec = dst->precision == MPP_PRECISION_FP64 ? sizeof(Rec<double>) : sizeof(Rec<float>);
    CUDA_TRY(cudaStreamSynchronize(src->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->d_mask, src->d_mask, sizeof(uint32_t) * dst->ncell, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->d_recs, src->d_recs, rec * (size_t)dst->ncell * MPP_CELL_CAPACITY, cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->d_nobj, src->d_nobj, sizeof(int), cudaMemcpyDeviceToDevice, dst->stream));
    CUDA_TRY(cudaMemcpyAsync(dst->d_next_uid, src->d_next_uid, sizeof(uint32_t), cudaMemcpyDeviceToDevice, dst->stream));
    return MPP_OK;
}

int mpp_pair_values(mpp_ctx *h, const uint32_t *handle_a, const uint32_t *handle_b, int n, double *out) {
    NEED(h, h->model_set, "mpp_pair_values: set the model first");
    if (n < 0 || (n > 0 && (!handle_a || !handle_b || !out))) return fail(MPP_ERR_INVALID, "mpp_pair_values: bad arguments");
    if (n == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    DISPATCH(h, (k_pair_values<R><<<blocks, WARPS_PER_BLOCK * 32, WARPS_PER_BLOCK * sizeof(Scratch<R>), h->stream>>>(
                    device_view<R>(h), handle_a, handle_b, n, out)));
    CUDA_TRY(cudaGetLastError());
    return check_device_errors(h);
}

int mpp_energy_vectors(mpp_ctx *h, const uint32_t *handle, int n, double *out_vectors, double *out_combined, double *out_totals) {
    NEED(h, h->model_set && (h->maps_set || h->m.setup == MPP_SETUP_TOY), "mpp_energy_vectors: set maps and model first");
    if (n < 0 || (n > 0 && !handle)) return fail(MPP_ERR_INVALID, "mpp_energy_vectors: bad arguments");
    if (out_totals && n > 0 && (!out_vectors || !out_combined)) return fail(MPP_ERR_INVALID, "mpp_energy_vectors: totals need both outputs");
    CUDA_TRY(cudaSetDevice(h->device));
    if (n > 0) {
        const int blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
        DISPATCH(h, (k_energy_vectors<R><<<blocks, WARPS_PER_BLOCK * 32, WARPS_PER_BLOCK * sizeof(Scratch<R>), h->stream>>>(
                        device_view<R>(h), handle, n, out_vectors, out_combined)));
    }
    if (out_totals) k_totals<<<1, 256, 0, h->stream>>>(out_vectors, out_combined, n, out_totals);
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

int mpp_delta_batch(mpp_ctx *h, const mpp_proposal *props, int m, double *out_delta) {
    NEED(h, h->model_set && (h->maps_set || h->m.setup == MPP_SETUP_TOY), "mpp_delta_batch: set maps and model first");
    if (m < 0 || (m > 0 && (!props || !out_delta))) return fail(MPP_ERR_INVALID, "mpp_delta_batch: bad arguments");
    if (m == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = (m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    DISPATCH(h, (k_delta_batch<R><<<blocks, WARPS_PER_BLOCK * 32, WARPS_PER_BLOCK * sizeof(Scratch<R>), h->stream>>>(
                    device_view<R>(h), props, m, out_delta)));
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

int mpp_replay(mpp_ctx *h, const mpp_proposal *props, int m, double t0, double alpha_t, double t_target, mpp_step_result *out) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_replay: set maps, model and kernels first");
    if (m < 0 || (m > 0 && (!props || !out)) || !(t0 > 0.0)) return fail(MPP_ERR_INVALID, "mpp_replay: bad arguments");
    if (m == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    DISPATCH(h, (k_replay<R><<<1, 32, sizeof(Scratch<R>), h->stream>>>(device_view<R>(h), props, m, t0, alpha_t, t_target, out)));
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

int mpp_run_sweeps(mpp_ctx *h, int n_sweeps, int per_visit, int stride, double t0, double alpha_t, double t_target,
                   uint64_t seed, uint64_t sweep_offset, unsigned long long *counters_host) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_run_sweeps: set maps, model and kernels first");
    if (n_sweeps < 0 || per_visit < 1 || stride < 3 || !(t0 > 0.0)) return fail(MPP_ERR_INVALID, "mpp_run_sweeps: bad arguments (stride >= 3)");
    CUDA_TRY(cudaSetDevice(h->device));
    double temp = t0;
    for (int s = 0; s < n_sweeps; ++s) {
        for (int ci = 0; ci < stride; ++ci)
            for (int cj = 0; cj < stride; ++cj) {
                const int n_ai = ci < h->nx ? (h->nx - ci + stride - 1) / stride : 0;
                const int n_aj = cj < h->ny ? (h->ny - cj + stride - 1) / stride : 0;
                const int n_active = n_ai * n_aj;
                if (n_active == 0) continue;
                const int blocks = (n_active + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
                DISPATCH(h, (k_sweep<R><<<blocks, WARPS_PER_BLOCK * 32, WARPS_PER_BLOCK * sizeof(Scratch<R>), h->stream>>>(
                                device_view<R>(h), stride, ci, cj, n_ai, n_aj, per_visit, temp, seed, sweep_offset + (uint64_t)s)));
            }
        if (temp > t_target) temp *= alpha_t;
    }
    CUDA_TRY(cudaGetLastError());
    if (counters_host) {
        unsigned long long *tmp = reinterpret_cast<unsigned long long *>(h->h_pinned);
        CUDA_TRY(cudaMemcpyAsync(tmp, h->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 8, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 8; ++i) counters_host[i] = tmp[i];
        return check_device_errors(h);
    }
    return MPP_OK;
}

int mpp_sample_births(mpp_ctx *h, int n, uint64_t seed, int32_t *out) {
    NEED(h, h->maps_set, "mpp_sample_births: set maps first");
    if (n < 0 || (n > 0 && !out)) return fail(MPP_ERR_INVALID, "mpp_sample_births: bad arguments");
    if (n == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    DISPATCH(h, (k_sample_births<R><<<blocks, WARPS_PER_BLOCK * 32, 0, h->stream>>>(device_view<R>(h), n, seed, out)));
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

int mpp_naive_init(mpp_ctx *h, double detection_threshold, double nms_distance, int *n_host) {
    NEED(h, h->maps_set && h->model_set, "mpp_naive_init: set maps and model first");
    if (h->map_rows != h->H) return fail(MPP_ERR_STATE, "mpp_naive_init: needs the maps of the whole scene (band-local maps are set)");
    if (nms_distance < 0 || nms_distance > 32) return fail(MPP_ERR_INVALID, "mpp_naive_init: nms_distance in [0,32]");
    CUDA_TRY(cudaSetDevice(h->device));
    const int n = h->H * h->W;
    if (!h->d_nms_state) CUDA_TRY(cudaMalloc(&h->d_nms_state, (size_t)n));
    const int rad = (int)floor(nms_distance), rad2 = (int)floor(nms_distance * nms_distance);
    const int blocks = (n + 255) / 256;
    int *flags = reinterpret_cast<int *>(h->d_counters);  // reuse: 2 ints
    k_nms_threshold<<<blocks, 256, 0, h->stream>>>(h->det, n, (float)detection_threshold, h->d_nms_state);
    // rounds of select / suppress are idempotent once nothing remains undecided, so they are issued six at a time and the host
    // only looks at the flags of the last one (one synchronisation per call on typical maps instead of one per round)
    for (int round = 0; round < 4096; ++round) {
        CUDA_TRY(cudaMemsetAsync(flags, 0, 2 * sizeof(int), h->stream));
        k_nms_select<<<blocks, 256, 0, h->stream>>>(h->det - (ptrdiff_t)h->map_row0 * h->W, h->H, h->W, rad, rad2, h->d_nms_state, flags);
        k_nms_suppress<<<blocks, 256, 0, h->stream>>>(h->H, h->W, rad, rad2, h->d_nms_state, flags + 1);
        if (round % 6 != 5) continue;
        CUDA_TRY(cudaMemcpyAsync(h->h_pinned, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (!h->h_pinned[1]) break;
        if (!h->h_pinned[0]) return fail(MPP_ERR_CUDA, "mpp_naive_init: NMS did not progress");
    }
    CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 8, h->stream));
    DISPATCH(h, (k_nms_insert<R><<<blocks, 256, 0, h->stream>>>(device_view<R>(h), h->d_nms_state)));
    CUDA_TRY(cudaGetLastError());
    int rc = check_device_errors(h);
    if (rc != MPP_OK) return rc;
    if (n_host) return mpp_num_objects(h, n_host);
    return MPP_OK;
}

int mpp_pack_rows(mpp_ctx *h, int row_lo, int row_hi, double *buf, int capacity, int *n_host) {
    NEED(h, true, "");
    if (capacity < 0 || (capacity > 0 && !buf) || !n_host) return fail(MPP_ERR_INVALID, "mpp_pack_rows: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    int *count = h->d_scan;  // scratch
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), h->stream));
    DISPATCH(h, (k_pack_rows<R><<<(h->ncell + 127) / 128, 128, 0, h->stream>>>(device_view<R>(h), row_lo, row_hi, buf, capacity, count)));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h->h_pinned, count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *n_host = *h->h_pinned;
    if (*n_host > capacity) return fail(MPP_ERR_INVALID, "mpp_pack_rows: buffer too small");
    return MPP_OK;
}

int mpp_unpack_rows(mpp_ctx *h, int row_lo, int row_hi, const double *buf, int n) {
    NEED(h, h->maps_set && h->model_set, "mpp_unpack_rows: set maps and model first");
    if (n < 0 || (n > 0 && !buf)) return fail(MPP_ERR_INVALID, "mpp_unpack_rows: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    DISPATCH(h, (k_erase_rows<R><<<(h->ncell + 127) / 128, 128, 0, h->stream>>>(device_view<R>(h), row_lo, row_hi)));
    if (n > 0) DISPATCH(h, (k_unpack_rows<R><<<(n + 127) / 128, 128, 0, h->stream>>>(device_view<R>(h), buf, n)));
    CUDA_TRY(cudaGetLastError());
    return check_device_errors(h);
}

}  // extern "C"

static int refresh_row_counts_impl(mpp_ctx *h) {
    CUDA_TRY(cudaMemsetAsync(h->d_nobj, 0, sizeof(int), h->stream));
    const int blocks = (h->nx + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    k_row_counts<<<blocks, WARPS_PER_BLOCK * 32, 0, h->stream>>>(h->d_mask, h->nx, h->ny, h->d_rowcount, h->d_nobj);
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

static int read_counters(mpp_ctx *h, unsigned long long *counters_host) {
    unsigned long long *tmp = reinterpret_cast<unsigned long long *>(h->h_pinned);
    CUDA_TRY(cudaMemcpyAsync(tmp, h->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 8, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 8; ++i) counters_host[i] = tmp[i];
    return check_device_errors(h);
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// uids of objects born in sweep s are 0x80000000 | ((s * cells + window) * 128 + proposal index): unique while the product stays
// below 2^31.  A call whose last sweep would exceed that is refused (two live objects could otherwise share a uid and the
// host mirror, which names objects by uid, would merge them).
static bool uid_space_ok(const mpp_ctx *h, uint64_t sweep_offset, int n_sweeps) {
    const uint64_t cells = (uint64_t)(h->nx + 2) * (uint64_t)(h->ny + 2);
    return (sweep_offset + (uint64_t)std::max(n_sweeps, 0)) * cells * (uint64_t)W2_PRE < 0x80000000ull;
}
#define UID_SPACE_MSG "sweep numbers beyond 2^31 / (128 * window count): the uids of born objects would wrap; restart the sweep numbering (sweep_offset) with another seed"

template <typename R, int NW, bool DBG, bool SIMT = false>
static cudaError_t launch_sweep2(mpp_ctx *h, int ci, int cj, int n_wi, int n_wj, int ox, int oy, int per_visit, float temp, uint64_t seed,
                                 uint64_t sweep_id, float *dbg) {
    const uint32_t uid_base = h->window_uid_next;
    h->window_uid_next += (uint32_t)(n_wi * n_wj * per_visit);
    if (h->window_uid_next < 0x80000000u) h->window_uid_next += 0x80000000u;  // wrapped: stay in the upper half
    const size_t smem = ((sizeof(WinState<R>) + 15) & ~(size_t)15) + (size_t)NW * W2_SCRATCH * sizeof(R);
    // the shared-memory opt-in is a per-device attribute of the function (a process may hold contexts on several GPUs);
    // setting it twice from two threads is harmless
    static bool configured[MPP_MAX_DEVICES] = {};
    if (!configured[h->device]) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep2<R, NW, DBG, SIMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[h->device] = true;
    }
    k_sweep2<R, NW, DBG, SIMT><<<n_wi * n_wj, 32 * NW, smem, h->stream>>>(device_view<R>(h), ci, cj, n_wi, n_wj, ox, oy, per_visit, temp, seed, sweep_id,
                                                                  uid_base, dbg);
    return cudaGetLastError();
}

template <typename R, int NW, bool DBG, bool SIMT = false>
static int launch_dataflow(mpp_ctx *h, int n_sweeps, int per_visit, double t0, double alpha_t, double t_target, uint64_t seed,
                           uint64_t sweep_offset, float *dbg) {
    const int S = n_sweeps;
    const int dg = (std::max(h->H, h->W) + 31) / 32 + 2;
    std::vector<int> ints((size_t)3 * S + 2);
    std::vector<float> temps(S);
    int *ox = ints.data(), *oy = ox + S, *base = oy + S;
    double temp = t0;
    int total = 0;
    for (int s = 0; s < S; ++s) {
        const uint64_t hsh = splitmix64(seed ^ splitmix64(sweep_offset + (uint64_t)s));
        ox[s] = (int)(hsh & 31); oy[s] = (int)((hsh >> 5) & 31);
        base[s] = total;
        total += ((h->H + ox[s] + 31) / 32) * ((h->W + oy[s] + 31) / 32);
        temps[s] = (float)temp;
        if (temp > t_target) temp *= alpha_t;
    }
    base[S] = total;
    // device layout: [ints: ox, oy, task_base | next_task | done 2*dg*dg] [floats: temp] [the device view of the context]
    const size_t n_int = (size_t)3 * S + 2 + 1 + (size_t)2 * dg * dg;
    const size_t off_ctx = (n_int * sizeof(int) + (size_t)S * sizeof(float) + 15) & ~(size_t)15;
    const size_t bytes = off_ctx + sizeof(Ctx<R>);
    if (bytes > h->plan_bytes) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_plan);
        h->d_plan = nullptr; h->plan_bytes = 0;
        CUDA_TRY(cudaMalloc(&h->d_plan, bytes * 2));
        h->plan_bytes = bytes * 2;
    }
    int *d_int = reinterpret_cast<int *>(h->d_plan);
    float *d_temp = reinterpret_cast<float *>(d_int + n_int);
    CUDA_TRY(cudaMemsetAsync(d_int + (size_t)3 * S + 2, 0, (1 + (size_t)2 * dg * dg) * sizeof(int), h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_int, ints.data(), ((size_t)3 * S + 2) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_temp, temps.data(), (size_t)S * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    const Ctx<R> view = device_view<R>(h);
    Ctx<R> *d_ctx = reinterpret_cast<Ctx<R> *>(reinterpret_cast<unsigned char *>(h->d_plan) + off_ctx);
    CUDA_TRY(cudaMemcpyAsync(d_ctx, &view, sizeof(Ctx<R>), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
    SweepPlan plan;
    plan.n_sweeps = S; plan.total_tasks = total; plan.dg = dg;
    plan.ox = d_int; plan.oy = d_int + S; plan.task_base = d_int + 2 * S;
    plan.next_task = d_int + 3 * S + 2; plan.done = d_int + 3 * S + 3; plan.temp = d_temp;
    const uint32_t uid_base = h->window_uid_next;
    h->window_uid_next += (uint32_t)total * (uint32_t)per_visit;
    if (h->window_uid_next < 0x80000000u) h->window_uid_next += 0x80000000u;
    const size_t smem = ((sizeof(WinState<R>) + 15) & ~(size_t)15) + (((size_t)NW * W2_SCRATCH * sizeof(R) + 15) & ~(size_t)15) + sizeof(Ctx<R>);
    static int blocks_per_sm_dev[MPP_MAX_DEVICES] = {};  // per device: shared-memory opt-in + occupancy of this instantiation
    if (!blocks_per_sm_dev[h->device]) {
        int bps = 0;
        CUDA_TRY(cudaFuncSetAttribute(k_windows_dataflow<R, NW, DBG, SIMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_windows_dataflow<R, NW, DBG, SIMT>, 32 * NW, smem));
        if (bps < 1) return fail(MPP_ERR_CUDA, "k_windows_dataflow does not fit on an SM");
        blocks_per_sm_dev[h->device] = bps;
    }
    const int blocks_per_sm = blocks_per_sm_dev[h->device];
    // persistent grid: never more CTAs than fit on the device (only CTAs that are running claim tasks, so the in-order
    // queue cannot deadlock), and not many more than can ever be active at once (about two colour classes of windows),
    // so that small scenes leave room for other contexts' kernels running concurrently on other streams (waiting CTAs
    // occupy SM slots: 32 tiles of 512^2 ran at 46 M proposals/s with two colour classes of CTAs each, 103 M/s with half a class)
    const int per_colour = (((h->H + 63) / 32 + 2) / 3) * (((h->W + 63) / 32 + 2) / 3);
    const int cap_x4 = SIMT ? 8 : 2;  // grid <= cap/4 colour classes + 8
    const int grid = std::max(1, std::min(std::min(total, blocks_per_sm * h->num_sms), cap_x4 * per_colour / 4 + 8));
    k_windows_dataflow<R, NW, DBG, SIMT><<<grid, 32 * NW, smem, h->stream>>>(view, d_ctx, plan, per_visit, seed, sweep_offset, uid_base, dbg);
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

extern "C" int mpp_run_windows(mpp_ctx *h, int n_sweeps, int per_visit, int n_warps, int schedule, double t0, double alpha_t, double t_target,
                               uint64_t seed, uint64_t sweep_offset, unsigned long long *counters_host, float *debug_maxdiff) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_run_windows: set maps, model and kernels first");
    if (n_sweeps < 0 || per_visit < 1 || per_visit > W2_PRE || !(t0 > 0.0)) return fail(MPP_ERR_INVALID, "mpp_run_windows: bad arguments (1 <= proposals_per_visit <= 128)");
    if (n_warps != 0 && n_warps != 1 && n_warps != 2 && n_warps != 4 && n_warps != 8)
        return fail(MPP_ERR_INVALID, "mpp_run_windows: n_warps must be 0 (lane-per-proposal mode), 1, 2, 4 or 8");
    if (schedule != 0 && schedule != 1) return fail(MPP_ERR_INVALID, "mpp_run_windows: schedule must be 0 (colour barriers) or 1 (dataflow)");
    if (h->m.setup == MPP_SETUP_TOY) return fail(MPP_ERR_STATE, "mpp_run_windows: needs a map-driven energy model");
    if (h->precision != MPP_PRECISION_FP32) return fail(MPP_ERR_STATE, "mpp_run_windows: the window sampler is float32 only (use mpp_run_chain / mpp_replay for float64)");
    if (!uid_space_ok(h, sweep_offset, n_sweeps)) return fail(MPP_ERR_INVALID, "mpp_run_windows: " UID_SPACE_MSG);
    CUDA_TRY(cudaSetDevice(h->device));
    // `alpha_t` is the temperature factor of one sweep; inside a visit the i-th proposal of every window stands for step
    // i * (number of windows) of the sweep, so the temperature decays by alpha_t^(1/per_visit) per proposal index and the
    // schedule is the reference's geometric one (rjmcmc.py:158-159) whatever the number of proposals per visit
    h->visit_alpha = (alpha_t > 0.0 && alpha_t < 1.0) ? (float)pow(alpha_t, 1.0 / (double)per_visit) : 1.f;
    h->visit_tfloor = (float)t_target;
    if (schedule == 1 && n_sweeps > 0) {
        int rc;
#define MPP_LAUNCH_D(NWV) (debug_maxdiff \
        ? launch_dataflow<float, NWV, true>(h, n_sweeps, per_visit, t0, alpha_t, t_target, seed, sweep_offset, debug_maxdiff) \
        : launch_dataflow<float, NWV, false>(h, n_sweeps, per_visit, t0, alpha_t, t_target, seed, sweep_offset, debug_maxdiff))
        switch (n_warps) {
        case 0: rc = debug_maxdiff ? launch_dataflow<float, 1, true, true>(h, n_sweeps, per_visit, t0, alpha_t, t_target, seed, sweep_offset, debug_maxdiff)
                                   : launch_dataflow<float, 1, false, true>(h, n_sweeps, per_visit, t0, alpha_t, t_target, seed, sweep_offset, debug_maxdiff);
                break;
        case 1: rc = MPP_LAUNCH_D(1); break;
        case 2: rc = MPP_LAUNCH_D(2); break;
        case 4: rc = MPP_LAUNCH_D(4); break;
        default: rc = MPP_LAUNCH_D(8); break;
        }
#undef MPP_LAUNCH_D
        if (rc != MPP_OK) return rc;
        if (counters_host) return read_counters(h, counters_host);
        return MPP_OK;
    }
    double temp = t0;
    for (int s = 0; s < n_sweeps; ++s) {
        const uint64_t sweep_id = sweep_offset + (uint64_t)s;
        const uint64_t hsh = splitmix64(seed ^ splitmix64(sweep_id));
        const int ox = (int)(hsh & 31), oy = (int)((hsh >> 5) & 31);
        const int nwx = (h->H + ox + 31) / 32, nwy = (h->W + oy + 31) / 32;
        for (int ci = 0; ci < 3; ++ci)
            for (int cj = 0; cj < 3; ++cj) {
                const int n_wi = ci < nwx ? (nwx - ci + 2) / 3 : 0, n_wj = cj < nwy ? (nwy - cj + 2) / 3 : 0;
                if (n_wi * n_wj == 0) continue;
                cudaError_t e;
#define MPP_LAUNCH_W(NWV) (debug_maxdiff \
        ? launch_sweep2<float, NWV, true>(h, ci, cj, n_wi, n_wj, ox, oy, per_visit, (float)temp, seed, sweep_id, debug_maxdiff) \
        : launch_sweep2<float, NWV, false>(h, ci, cj, n_wi, n_wj, ox, oy, per_visit, (float)temp, seed, sweep_id, debug_maxdiff))
                switch (n_warps) {
                case 0: e = debug_maxdiff ? launch_sweep2<float, 1, true, true>(h, ci, cj, n_wi, n_wj, ox, oy, per_visit, (float)temp, seed, sweep_id, debug_maxdiff)
                                          : launch_sweep2<float, 1, false, true>(h, ci, cj, n_wi, n_wj, ox, oy, per_visit, (float)temp, seed, sweep_id, debug_maxdiff);
                        break;
                case 1: e = MPP_LAUNCH_W(1); break;
                case 2: e = MPP_LAUNCH_W(2); break;
                case 4: e = MPP_LAUNCH_W(4); break;
                default: e = MPP_LAUNCH_W(8); break;
                }
#undef MPP_LAUNCH_W
                if (e != cudaSuccess) return fail(MPP_ERR_CUDA, std::string("k_sweep2 launch: ") + cudaGetErrorString(e));
            }
        if (temp > t_target) temp *= alpha_t;
    }
    if (counters_host) return read_counters(h, counters_host);
    return MPP_OK;
}

extern "C" int mpp_run_window_rows(mpp_ctx *h, int per_visit, int n_warps, double temperature, uint64_t seed, uint64_t sweep_id, int ci,
                                   int row_lo, int row_hi) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_run_window_rows: set maps, model and kernels first");
    if (per_visit < 1 || per_visit > W2_PRE || !(temperature > 0.0) || ci < 0 || ci > 2 || row_lo > row_hi)
        return fail(MPP_ERR_INVALID, "mpp_run_window_rows: bad arguments");
    if (n_warps != 0 && n_warps != 1 && n_warps != 2 && n_warps != 4 && n_warps != 8) return fail(MPP_ERR_INVALID, "mpp_run_window_rows: n_warps must be 0, 1, 2, 4 or 8");
    if (h->m.setup == MPP_SETUP_TOY || h->precision != MPP_PRECISION_FP32) return fail(MPP_ERR_STATE, "mpp_run_window_rows: float32 map-driven model only");
    if (!uid_space_ok(h, sweep_id, 1)) return fail(MPP_ERR_INVALID, "mpp_run_window_rows: " UID_SPACE_MSG);
    CUDA_TRY(cudaSetDevice(h->device));
    h->visit_alpha = 1.f; h->visit_tfloor = 0.f;  // one temperature per phase
    const uint64_t hsh = splitmix64(seed ^ splitmix64(sweep_id));
    const int ox = (int)(hsh & 31), oy = (int)((hsh >> 5) & 31);
    const int nwx = (h->H + ox + 31) / 32, nwy = (h->W + oy + 31) / 32;
    // window rows whose first pixel row max(32*wi - ox, 0) lies in [row_lo, row_hi) and wi = ci (mod 3)
    int first = -1, count = 0;
    for (int wi = ci; wi < nwx; wi += 3) {
        const int start = std::max(32 * wi - ox, 0);
        if (start >= row_lo && start < row_hi) { if (first < 0) first = wi; ++count; }
    }
    if (count == 0) return MPP_OK;
    for (int cj = 0; cj < 3; ++cj) {
        const int n_wj = cj < nwy ? (nwy - cj + 2) / 3 : 0;
        if (n_wj == 0) continue;
        cudaError_t e;
        switch (n_warps) {
        case 0: e = launch_sweep2<float, 1, false, true>(h, first, cj, count, n_wj, ox, oy, per_visit, (float)temperature, seed, sweep_id, nullptr); break;
        case 1: e = launch_sweep2<float, 1, false>(h, first, cj, count, n_wj, ox, oy, per_visit, (float)temperature, seed, sweep_id, nullptr); break;
        case 2: e = launch_sweep2<float, 2, false>(h, first, cj, count, n_wj, ox, oy, per_visit, (float)temperature, seed, sweep_id, nullptr); break;
        case 4: e = launch_sweep2<float, 4, false>(h, first, cj, count, n_wj, ox, oy, per_visit, (float)temperature, seed, sweep_id, nullptr); break;
        default: e = launch_sweep2<float, 8, false>(h, first, cj, count, n_wj, ox, oy, per_visit, (float)temperature, seed, sweep_id, nullptr); break;
        }
        if (e != cudaSuccess) return fail(MPP_ERR_CUDA, std::string("k_sweep2 launch: ") + cudaGetErrorString(e));
    }
    return MPP_OK;
}

extern "C" int mpp_window_stats(mpp_ctx *h, unsigned long long *out_host) {
    if (!h || !out_host) return fail(MPP_ERR_INVALID, "mpp_window_stats: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<unsigned long long> tmp(MPP_KSTATS_ALLOC);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->d_kstats, sizeof(unsigned long long) * MPP_KSTATS_ALLOC, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d_kstats, 0, sizeof(unsigned long long) * MPP_KSTATS_ALLOC, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < MPP_KSTATS_COPY; ++i) out_host[i] = tmp[i];
    return MPP_OK;
}

extern "C" int mpp_set_window_trace(mpp_ctx *h, mpp_window_trace *buf, uint64_t capacity, uint64_t sweep0) {
    if (!h) return fail(MPP_ERR_INVALID, "mpp_set_window_trace: null context");
    h->trace = buf; h->trace_capacity = buf ? capacity : 0; h->trace_sweep0 = sweep0;
    return MPP_OK;
}

// ================================================================================================ multi-scene / split-scene schedule
static int done_grid_pitch(const mpp_ctx *h) { return (std::max(h->H, h->W) + 31) / 32 + 2; }

static int ensure_done_grid(mpp_ctx *h) {
    if (h->d_done) return MPP_OK;
    const int dg = done_grid_pitch(h);
    CUDA_TRY(cudaMalloc(&h->d_done, sizeof(int) * 2 * dg * dg));
    CUDA_TRY(cudaMemset(h->d_done, 0, sizeof(int) * 2 * dg * dg));
    h->sweeps_done = 0;
    return MPP_OK;
}

template <int NW, bool DBG, bool SPLIT>
static int launch_multi(mpp_ctx **ctxs, const uint64_t *seeds, int n, uint64_t grid_seed, int n_sweeps, int per_visit, double t0, double alpha_t,
                        double t_target, uint64_t sweep_offset, int max_ctas, float *dbg) {
    typedef float R;
    mpp_ctx *h0 = ctxs[0];
    const int S = n_sweeps, H = h0->H, W = h0->W, dg = done_grid_pitch(h0);
    std::vector<int> ints((size_t)4 * S + 2 + 9 * S + 1);
    int *ox = ints.data(), *oy = ox + S + 1, *wlo = oy + S + 1, *whi = wlo + S, *base = whi + S;
    std::vector<float> temps(S);
    ox[0] = h0->last_ox; oy[0] = h0->last_oy;
    double temp = t0;
    long long total = 0;
    for (int s = 0; s < S; ++s) {
        const uint64_t hsh = splitmix64(grid_seed ^ splitmix64(sweep_offset + (uint64_t)s));
        ox[s + 1] = (int)(hsh & 31); oy[s + 1] = (int)((hsh >> 5) & 31);
        const int nwx = (H + ox[s + 1] + 31) / 32, nwy = (W + oy[s + 1] + 31) / 32;
        // window rows of this rank: first pixel row max(32 wi - ox, 0) in [row_lo, row_hi)
        wlo[s] = 0; whi[s] = nwx;
        if (SPLIT) {
            wlo[s] = h0->row_lo <= 0 ? 0 : (h0->row_lo + ox[s + 1] + 31) / 32;
            whi[s] = h0->row_hi >= H ? nwx : std::min(nwx, (h0->row_hi + ox[s + 1] + 31) / 32);
        }
        for (int col = 0; col < 9; ++col) {
            const int ci = col / 3, cj = col % 3;
            const int first_i = wlo[s] + (ci - wlo[s] % 3 + 3) % 3;
            const int a_i = first_i < whi[s] ? (whi[s] - first_i + 2) / 3 : 0, a_j = cj < nwy ? (nwy - cj + 2) / 3 : 0;
            base[9 * s + col] = (int)total;
            total += (long long)n * a_i * a_j;
        }
        temps[s] = (float)temp;
        if (temp > t_target) temp *= alpha_t;
    }
    base[9 * S] = (int)total;
    if (total > 0x7fffffffLL) return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: more than 2^31 window visits in one call");
    // scene descriptors
    std::vector<SceneDev<R>> scenes(n);
    for (int k = 0; k < n; ++k) {
        mpp_ctx *h = ctxs[k];
        SceneDev<R> &d = scenes[k];
        memset(&d, 0, sizeof(d));
        d.c = device_view<R>(h);
        d.seed = seeds[k];
        d.done = h->d_done; d.done_up = h->done_up; d.done_down = h->done_down;
        d.notify_lo = h->row_lo + 160; d.notify_hi = h->row_hi - 160;
    }
    // device layout in the plan scratch of the first context: [ints | next_task | temps | scenes]
    const size_t n_int = ints.size() + 1;
    size_t off_temp = n_int * sizeof(int), off_scenes = (off_temp + (size_t)S * sizeof(float) + 15) & ~(size_t)15;
    const size_t bytes = off_scenes + (size_t)n * sizeof(SceneDev<R>);
    if (bytes > h0->plan_bytes) {
        CUDA_TRY(cudaStreamSynchronize(h0->stream));
        cudaFree(h0->d_plan);
        h0->d_plan = nullptr; h0->plan_bytes = 0;
        CUDA_TRY(cudaMalloc(&h0->d_plan, bytes * 2));
        h0->plan_bytes = bytes * 2;
    }
    unsigned char *d_base = reinterpret_cast<unsigned char *>(h0->d_plan);
    int *d_int = reinterpret_cast<int *>(d_base);
    CUDA_TRY(cudaMemcpyAsync(d_int, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice, h0->stream));
    CUDA_TRY(cudaMemsetAsync(d_int + ints.size(), 0, sizeof(int), h0->stream));
    CUDA_TRY(cudaMemcpyAsync(d_base + off_temp, temps.data(), (size_t)S * sizeof(float), cudaMemcpyHostToDevice, h0->stream));
    CUDA_TRY(cudaMemcpyAsync(d_base + off_scenes, scenes.data(), (size_t)n * sizeof(SceneDev<R>), cudaMemcpyHostToDevice, h0->stream));
    CUDA_TRY(cudaStreamSynchronize(h0->stream));  // the host vectors go out of scope
    MultiPlan plan;
    plan.n_sweeps = S; plan.n_scenes = n; plan.total_tasks = (int)total; plan.dg = dg;
    plan.stamp0 = h0->sweeps_done + 1;
    plan.ox = d_int; plan.oy = d_int + S + 1; plan.wi_lo = d_int + 2 * S + 2; plan.wi_hi = d_int + 3 * S + 2; plan.task_base = d_int + 4 * S + 2;
    plan.next_task = d_int + ints.size();
    plan.temp = reinterpret_cast<const float *>(d_base + off_temp);
    constexpr size_t WS = (sizeof(WinState<R>) + 15) & ~(size_t)15, SC = ((size_t)NW * W2_SCRATCH * sizeof(R) + 15) & ~(size_t)15;
    const size_t smem = WS + SC + sizeof(SceneDev<R>);
    static int blocks_per_sm_dev[MPP_MAX_DEVICES] = {};
    if (!blocks_per_sm_dev[h0->device]) {
        int bps = 0;
        CUDA_TRY(cudaFuncSetAttribute(k_windows_multi<R, NW, DBG, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_windows_multi<R, NW, DBG, SPLIT>, 32 * NW, smem));
        if (bps < 1) return fail(MPP_ERR_CUDA, "k_windows_multi does not fit on an SM");
        blocks_per_sm_dev[h0->device] = bps;
    }
    // persistent grid: never more CTAs than are resident at once (only running CTAs claim visits); about half a colour class of
    // every scene can be active at a time
    const int rows = SPLIT ? (h0->row_hi - h0->row_lo) : H;
    const long long per_colour = (long long)n * (((rows + 63) / 32 + 2) / 3) * (((W + 63) / 32 + 2) / 3);
    long long grid = std::min<long long>(std::min<long long>(total, (long long)blocks_per_sm_dev[h0->device] * h0->num_sms), per_colour / 2 + 8);
    if (max_ctas > 0) grid = std::min<long long>(grid, max_ctas);
    grid = std::max<long long>(grid, 1);
    if (total > 0)
        k_windows_multi<R, NW, DBG, SPLIT><<<(int)grid, 32 * NW, smem, h0->stream>>>(scenes[0], reinterpret_cast<const SceneDev<R> *>(d_base + off_scenes), plan, per_visit,
                                                                                    sweep_offset, dbg);
    CUDA_TRY(cudaGetLastError());
    for (int k = 0; k < n; ++k) {
        ctxs[k]->sweeps_done += S;
        if (S > 0) { ctxs[k]->last_ox = ox[S]; ctxs[k]->last_oy = oy[S]; }
    }
    return MPP_OK;
}

extern "C" int mpp_run_windows_batch(mpp_ctx **ctxs, const uint64_t *seeds, int n_scenes, uint64_t grid_seed, int n_sweeps, int per_visit,
                                     int n_warps, double t0, double alpha_t, double t_target, uint64_t sweep_offset, int max_ctas,
                                     unsigned long long *counters_host, float *debug_maxdiff) {
    if (!ctxs || !seeds || n_scenes < 1) return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: no scenes");
    if (n_sweeps < 0 || per_visit < 1 || per_visit > W2_PRE || !(t0 > 0.0)) return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: bad arguments (1 <= proposals_per_visit <= 128)");
    if (n_warps != 4 && n_warps != 8) return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: n_warps must be 4 or 8");
    mpp_ctx *h0 = ctxs[0];
    for (int k = 0; k < n_scenes; ++k) {
        mpp_ctx *h = ctxs[k];
        NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_run_windows_batch: set maps, model and kernels of every scene first");
        if (h->H != h0->H || h->W != h0->W || h->device != h0->device || h->precision != MPP_PRECISION_FP32 || h->m.setup == MPP_SETUP_TOY)
            return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: scenes must share shape and device, float32 map-driven models only");
        if (h->split && n_scenes != 1) return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: a split scene runs alone on its rank");
    }
    if (!uid_space_ok(h0, sweep_offset, n_sweeps)) return fail(MPP_ERR_INVALID, "mpp_run_windows_batch: " UID_SPACE_MSG);
    CUDA_TRY(cudaSetDevice(h0->device));
    for (int k = 0; k < n_scenes; ++k) {
        mpp_ctx *h = ctxs[k];
        const int rc = ensure_done_grid(h);
        if (rc != MPP_OK) return rc;
        if (h->sweeps_done != h0->sweeps_done || h->last_ox != h0->last_ox || h->last_oy != h0->last_oy)
            return fail(MPP_ERR_STATE, "mpp_run_windows_batch: the scenes of a batch must have run the same sweeps before (reset them together)");
        h->visit_alpha = (alpha_t > 0.0 && alpha_t < 1.0) ? (float)pow(alpha_t, 1.0 / (double)per_visit) : 1.f;
        h->visit_tfloor = (float)t_target;
        if (h->stream != h0->stream) CUDA_TRY(cudaStreamSynchronize(h->stream));  // its maps / objects were set up on its own stream
    }
    int rc;
#define MPP_LAUNCH_M(NWV, SPL) (debug_maxdiff \
        ? launch_multi<NWV, true, SPL>(ctxs, seeds, n_scenes, grid_seed, n_sweeps, per_visit, t0, alpha_t, t_target, sweep_offset, max_ctas, debug_maxdiff) \
        : launch_multi<NWV, false, SPL>(ctxs, seeds, n_scenes, grid_seed, n_sweeps, per_visit, t0, alpha_t, t_target, sweep_offset, max_ctas, debug_maxdiff))
    if (h0->split) rc = n_warps == 4 ? MPP_LAUNCH_M(4, true) : MPP_LAUNCH_M(8, true);
    else rc = n_warps == 4 ? MPP_LAUNCH_M(4, false) : MPP_LAUNCH_M(8, false);
#undef MPP_LAUNCH_M
    if (rc != MPP_OK) return rc;
    if (counters_host) {
        for (int i = 0; i < 8; ++i) counters_host[i] = 0;
        std::vector<unsigned long long> tmp((size_t)8 * n_scenes);
        std::vector<uint32_t> errs(n_scenes);
        for (int k = 0; k < n_scenes; ++k) {
            CUDA_TRY(cudaMemcpyAsync(tmp.data() + 8 * k, ctxs[k]->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, h0->stream));
            CUDA_TRY(cudaMemsetAsync(ctxs[k]->d_counters, 0, sizeof(unsigned long long) * 8, h0->stream));
            CUDA_TRY(cudaMemcpyAsync(&errs[k], ctxs[k]->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, h0->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(h0->stream));
        for (int k = 0; k < n_scenes; ++k) {
            for (int i = 0; i < 8; ++i) counters_host[i] += tmp[8 * k + i];
            if (errs[k]) { const int e = check_device_errors(ctxs[k]); if (e != MPP_OK) return e; }
        }
    }
    return MPP_OK;
}

// ---- scene split across GPUs: peer mappings of the neighbours' state
extern "C" int mpp_split_export(mpp_ctx *h, unsigned char *handles_host) {
    if (!h || !handles_host) return fail(MPP_ERR_INVALID, "mpp_split_export: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const int rc = ensure_done_grid(h);
    if (rc != MPP_OK) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == MPP_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    void *ptrs[3] = {h->d_mask, h->d_recs, h->d_done};
    for (int i = 0; i < 3; ++i) {
        cudaIpcMemHandle_t hd;
        CUDA_TRY(cudaIpcGetMemHandle(&hd, ptrs[i]));
        memcpy(handles_host + (size_t)i * MPP_IPC_HANDLE_BYTES, &hd, MPP_IPC_HANDLE_BYTES);
    }
    return MPP_OK;
}

static int split_common(mpp_ctx *h, int row_lo, int row_hi) {
    if (row_lo < 0 || row_hi > h->H || row_lo >= row_hi || row_lo % MPP_CELL_SIZE || (row_hi % MPP_CELL_SIZE && row_hi != h->H))
        return fail(MPP_ERR_INVALID, "mpp_split_attach: the band must be a range of whole 32-px cell rows");
    if ((row_lo > 0 || row_hi < h->H) && row_hi - row_lo < 384)
        return fail(MPP_ERR_INVALID, "mpp_split_attach: bands must be at least 384 rows (a window only ever reaches its two neighbour bands)");
    if (h->precision != MPP_PRECISION_FP32) return fail(MPP_ERR_STATE, "mpp_split_attach: float32 contexts only");
    const int rc = ensure_done_grid(h);
    if (rc != MPP_OK) return rc;
    // the plan scratch is allocated now: cudaFree / cudaMalloc at run time would synchronise the device while a neighbour band's
    // kernel (same device, in the tests) is waiting for this band to start
    const size_t want = (size_t)1 << 20;
    if (h->plan_bytes < want) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_plan);
        h->d_plan = nullptr; h->plan_bytes = 0;
        CUDA_TRY(cudaMalloc(&h->d_plan, want));
        h->plan_bytes = want;
    }
    h->split = true; h->row_lo = row_lo; h->row_hi = row_hi;
    return MPP_OK;
}

extern "C" int mpp_split_attach(mpp_ctx *h, int row_lo, int row_hi, const unsigned char *up_handles_host, const unsigned char *down_handles_host) {
    if (!h) return fail(MPP_ERR_INVALID, "mpp_split_attach: null context");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = split_common(h, row_lo, row_hi);
    if (rc != MPP_OK) return rc;
    const unsigned char *src[2] = {up_handles_host, down_handles_host};
    for (int side = 0; side < 2; ++side) {
        void *opened[3] = {nullptr, nullptr, nullptr};
        if (src[side])
            for (int i = 0; i < 3; ++i) {
                cudaIpcMemHandle_t hd;
                memcpy(&hd, src[side] + (size_t)i * MPP_IPC_HANDLE_BYTES, MPP_IPC_HANDLE_BYTES);
                CUDA_TRY(cudaIpcOpenMemHandle(&opened[i], hd, cudaIpcMemLazyEnablePeerAccess));
                h->ipc_opened[3 * side + i] = opened[i];
            }
        if (side == 0) { h->mask_up = (uint32_t *)opened[0]; h->recs_up = opened[1]; h->done_up = (int *)opened[2]; }
        else { h->mask_down = (uint32_t *)opened[0]; h->recs_down = opened[1]; h->done_down = (int *)opened[2]; }
    }
    return MPP_OK;
}

extern "C" int mpp_split_attach_local(mpp_ctx *h, int row_lo, int row_hi, mpp_ctx *up, mpp_ctx *down) {
    if (!h) return fail(MPP_ERR_INVALID, "mpp_split_attach_local: null context");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = split_common(h, row_lo, row_hi);
    if (rc != MPP_OK) return rc;
    mpp_ctx *nb[2] = {up, down};
    for (int side = 0; side < 2; ++side) {
        mpp_ctx *o = nb[side];
        if (o) {
            if (o->H != h->H || o->W != h->W || o->precision != h->precision) return fail(MPP_ERR_INVALID, "mpp_split_attach_local: neighbour of another shape");
            rc = ensure_done_grid(o);
            if (rc != MPP_OK) return rc;
            if (o->device != h->device) {
                int can = 0;
                CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->device, o->device));
                if (!can) return fail(MPP_ERR_CUDA, "mpp_split_attach_local: no peer access between the two devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(MPP_ERR_CUDA, cudaGetErrorString(e));
                cudaGetLastError();
                CUDA_TRY(cudaSetDevice(h->device));
            }
        }
        if (side == 0) { h->mask_up = o ? o->d_mask : nullptr; h->recs_up = o ? o->d_recs : nullptr; h->done_up = o ? o->d_done : nullptr; }
        else { h->mask_down = o ? o->d_mask : nullptr; h->recs_down = o ? o->d_recs : nullptr; h->done_down = o ? o->d_done : nullptr; }
    }
    return MPP_OK;
}

extern "C" int mpp_split_detach(mpp_ctx *h) {
    if (!h) return fail(MPP_ERR_INVALID, "mpp_split_detach: null context");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (void *&p : h->ipc_opened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    h->split = false; h->row_lo = 0; h->row_hi = 0;
    h->mask_up = h->mask_down = nullptr; h->recs_up = h->recs_down = nullptr; h->done_up = h->done_down = nullptr;
    return MPP_OK;
}

extern "C" int mpp_window_grid(mpp_ctx *h, uint64_t seed, uint64_t sweep_id, int *ox_host, int *oy_host) {
    if (!h || !ox_host || !oy_host) return fail(MPP_ERR_INVALID, "mpp_window_grid: null argument");
    const uint64_t hsh = splitmix64(seed ^ splitmix64(sweep_id));
    *ox_host = (int)(hsh & 31); *oy_host = (int)((hsh >> 5) & 31);
    return MPP_OK;
}

extern "C" int mpp_run_chain(mpp_ctx *h, int n_steps, double t0, double alpha_t, double t_target, uint64_t seed, uint64_t step_offset,
                  mpp_step_result *trace, unsigned long long *counters_host) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_run_chain: set maps, model and kernels first");
    if (h->map_rows != h->H) return fail(MPP_ERR_STATE, "mpp_run_chain: the global kernels need the maps of the whole scene (band-local maps are set)");
    if (n_steps < 0 || !(t0 > 0.0)) return fail(MPP_ERR_INVALID, "mpp_run_chain: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    if (n_steps > 0) {
        int rc = refresh_row_counts(h);
        if (rc != MPP_OK) return rc;
        DISPATCH(h, (k_run_chain<R><<<1, 32, sizeof(Scratch<R>), h->stream>>>(device_view<R>(h), h->d_rowcount, n_steps, t0, alpha_t,
                                                                           t_target, seed, step_offset, trace)));
        CUDA_TRY(cudaGetLastError());
    }
    if (counters_host) return read_counters(h, counters_host);
    return MPP_OK;
}

extern "C" int mpp_sample_proposals(mpp_ctx *h, const int32_t *kernel_ids, int m, uint64_t seed, uint64_t offset, mpp_proposal *out) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_sample_proposals: set maps, model and kernels first");
    if (m < 0 || (m > 0 && !out)) return fail(MPP_ERR_INVALID, "mpp_sample_proposals: bad arguments");
    if (m == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = refresh_row_counts(h);
    if (rc != MPP_OK) return rc;
    const int blocks = (m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    DISPATCH(h, (k_sample_proposals<R><<<blocks, WARPS_PER_BLOCK * 32, 0, h->stream>>>(device_view<R>(h), h->d_rowcount, kernel_ids, m,
                                                                                    seed, offset, out)));
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

extern "C" int mpp_sample_split_merge(mpp_ctx *h, int kind, double radius, const double *sig, uint64_t seed, uint64_t offset, mpp_split_merge *out) {
    NEED(h, h->model_set && h->kernels_set, "mpp_sample_split_merge: set the model and the kernels first");
    if ((kind != 8 && kind != 9) || !(radius > 0.0) || radius > 32.0 || !sig || !out) return fail(MPP_ERR_INVALID, "mpp_sample_split_merge: bad arguments (radius <= 32)");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = refresh_row_counts(h);
    if (rc != MPP_OK) return rc;
    DISPATCH(h, (k_sample_split_merge<R><<<1, 32, 0, h->stream>>>(device_view<R>(h), h->d_rowcount, kind, radius, sig[0], sig[1], sig[2], seed, offset, out)));
    CUDA_TRY(cudaGetLastError());
    return check_device_errors(h);
}

extern "C" int mpp_split_merge_probs(mpp_ctx *h, const mpp_split_merge *p, double p_split, double p_merge, double radius, const double *sig,
                                     double *out) {
    NEED(h, h->model_set && h->kernels_set, "mpp_split_merge_probs: set the model and the kernels first");
    if (!p || !sig || !out || !(radius > 0.0) || radius > 32.0) return fail(MPP_ERR_INVALID, "mpp_split_merge_probs: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = refresh_row_counts(h);  // (recounts the objects: the densities use len(x))
    if (rc != MPP_OK) return rc;
    DISPATCH(h, (k_split_merge_probs<R><<<1, 32, 0, h->stream>>>(device_view<R>(h), p, p_split, p_merge, radius, sig[0], sig[1], sig[2], out)));
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

extern "C" int mpp_proposal_probs(mpp_ctx *h, const mpp_proposal *props, int m, double *out) {
    NEED(h, h->maps_set && h->model_set && h->kernels_set, "mpp_proposal_probs: set maps, model and kernels first");
    if (m < 0 || (m > 0 && (!props || !out))) return fail(MPP_ERR_INVALID, "mpp_proposal_probs: bad arguments");
    if (m == 0) return MPP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = (m + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    DISPATCH(h, (k_proposal_probs<R><<<blocks, WARPS_PER_BLOCK * 32, 0, h->stream>>>(device_view<R>(h), props, m, out)));
    CUDA_TRY(cudaGetLastError());
    return check_device_errors(h);
}

extern "C" int mpp_combine(const mpp_model_params *model_host, const double *vectors, int n, double *out_per_object, double *out_total,
                int device, void *stream) {
    if (n < 0 || (n > 0 && !vectors) || !out_total) return fail(MPP_ERR_INVALID, "mpp_combine: bad arguments");
    ModelDev m;
    memset(&m, 0, sizeof(m));
    const int rc = model_from_params(model_host, m);
    if (rc != MPP_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    k_combine<<<1, 256, 0, (cudaStream_t)stream>>>(m, vectors, n, out_per_object, out_total);
    CUDA_TRY(cudaGetLastError());
    return MPP_OK;
}

