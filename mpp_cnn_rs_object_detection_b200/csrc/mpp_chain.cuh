// Proposal construction shared by the sequential device chain (reference semantics: global kernels of
// make_kernels.py:88-144) and the proposal-sampling entry point.  Rows R15-R21 of SURVEY.md section 8a.
#pragma once
#include "mpp_device.cuh"
#include "mpp_proposals.cuh"

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, uint32_t cc, uint32_t d, double *n0, double *n1) {
    const double u1 = u01(a, b), u2 = u01(cc, d);
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    *n0 = r * cs; *n1 = r * sn;
}

__device__ __forceinline__ int draw_kernel(const KernDev &k, double u) {  // rng.choice(kernels, p) rjmcmc.py:88
    double acc = 0, tot = 0;
    for (int i = 0; i < 8; ++i) tot += k.p[i];
    const double t = u * tot;
    for (int i = 0; i < 7; ++i) { acc += k.p[i]; if (t < acc) return i; }
    return 7;
}

// UniformRectangleSampler.sample (shape_samplers.py:136-141) inside the pixel box [x0,x1) x [y0,y1)
template <typename R>
__device__ __forceinline__ Rec<R> uniform_birth(const Ctx<R> &c, int x0, int x1, int y0, int y1, float ux, float uy,
                                                double us, double ur, double ua) {
    const int wx = x1 - x0, wy = y1 - y0;
    const int x = x0 + min(wx - 1, (int)(ux * (float)wx)), y = y0 + min(wy - 1, (int)(uy * (float)wy));
    const R size = (R)(us * 32.0), ratio = (R)ur, angle = (R)(ua * 3.14159265358979323846);
    return make_rec(c, x, y, size, ratio, angle,
                    pack_cls(value_to_class<R>(0, size), value_to_class<R>(1, ratio), value_to_class<R>(2, angle)), 0u);
}

// marks of a data-driven birth at pixel (x, y): RectangleSampler._sample_param (shape_samplers.py:113-117)
template <typename R>
__device__ __forceinline__ Rec<R> data_birth_at(const Ctx<R> &c, int x, int y, float u0, float u1, float u2, int lane) {
    const int c0 = sample_mark_class(c, 0, x, y, u0, lane);
    const int c1 = sample_mark_class(c, 1, x, y, u1, lane);
    const int c2 = sample_mark_class(c, 2, x, y, u2, lane);
    return make_rec(c, x, y, mark_edge<R>(0, c0), mark_edge<R>(1, c1), mark_edge<R>(2, c2), pack_cls(c0, c1, c2), 0u);
}

// Kernels 4..7 applied to the picked object `rem` (transform_kernels.py).  r1 = four random words, u_param in [0,1).
// The unit energies of `add` are NOT refreshed here (callers do it once the move is known to be admissible).
template <typename R>
__device__ __forceinline__ void propose_move(const Ctx<R> &c, int kernel, const Rec<R> &rem, uint4 r1, float u_param, int lane,
                                             Rec<R> *add, double *d0, double *d1, int *param_id, int *new_class) {
    *add = rem;
    *d0 = 0; *d1 = 0; *param_id = 0; *new_class = 0;
    switch (kernel) {
    case 4: {  // GaussianTranslationKernel.sample_perturbation transform_kernels.py:24-36
        box_muller(r1.x, r1.y, r1.z, r1.w, d0, d1);
        *d0 *= c.k.trl_sigma; *d1 *= c.k.trl_sigma;
        int nx_ = (int)((double)rem.x + *d0), ny_ = (int)((double)rem.y + *d1);  // astype(int): truncation
        add->x = min(max(nx_, 0), c.H - 1); add->y = min(max(ny_, 0), c.W - 1);
        return;
    }
    case 5: {  // DataDrivenTranslationKernel.sample_perturbation :77-89
        const int md = c.k.trl_max_delta;
        int x, y;
        sample_window(c, max(0, rem.x - md), min(rem.x + md + 1, c.H), max(0, rem.y - md), min(rem.y + md + 1, c.W),
                      u01f(r1.x), u01f(r1.y), lane, &x, &y, nullptr);
        add->x = x; add->y = y;
        return;
    }
    case 6: {  // GaussianShapeTransformKernel.sample_perturbation :128-143
        const int pid = min(2, (int)(u_param * 3.0f));
        double dummy;
        box_muller(r1.x, r1.y, r1.z, r1.w, d0, &dummy);
        *d0 *= c.k.trf_sigma[pid];
        R v = (pid == 0 ? rem.size : (pid == 1 ? rem.ratio : rem.angle)) + (R)*d0;
        const R vmax = (R)mark_vmax(pid);
        if (pid == 2) { v = v - r_floor(v / vmax) * vmax; if (!(v < vmax)) v = 0; if (v < 0) v = 0; }  // python % on a cyclic mark
        else v = r_min(r_max(v, (R)0), vmax);
        if (pid == 0) add->size = v; else if (pid == 1) add->ratio = v; else add->angle = v;
        const int nc = value_to_class<R>(pid, v);
        add->cls = (rem.cls & ~(0xffu << (8 * pid))) | ((uint32_t)nc << (8 * pid));
        fill_geometry(*add);
        *param_id = pid;
        return;
    }
    default: {  // DataDrivenShapeTransformKernel.sample_perturbation :179-200
        const int pid = min(2, (int)(u_param * 3.0f));
        const int ncls = sample_mark_class(c, pid, rem.x, rem.y, u01f(r1.x), lane);
        const R v = mark_edge<R>(pid, ncls);
        if (pid == 0) add->size = v; else if (pid == 1) add->ratio = v; else add->angle = v;
        add->cls = (rem.cls & ~(0xffu << (8 * pid))) | ((uint32_t)ncls << (8 * pid));
        fill_geometry(*add);
        *param_id = pid; *new_class = ncls;
        return;
    }
    }
}

// PointsSet.random_choice (point_set.py:151-185): the i-th object in cell-major order, i uniform in [0, n).
// row_count [nx] = objects per row of cells.  Returns MPP_NO_OBJECT if the counts are inconsistent.
template <typename R>
__device__ __forceinline__ uint32_t pick_global(const Ctx<R> &c, const int *row_count, int n, double u, int lane) {
    int target = min(n - 1, (int)(u * (double)n));
    int row = -1;
    for (int b = 0; b < c.nx && row < 0; b += 32) {
        const int cnt = (b + lane < c.nx) ? __ldcg(row_count + b + lane) : 0;
        const int incl = warp_incl_scan(cnt, lane);
        const int tot = __shfl_sync(MPP_FULL, incl, 31);
        if (target < tot) {
            const uint32_t bal = __ballot_sync(MPP_FULL, incl > target);
            const int l = __ffs(bal) - 1;
            row = b + l;
            target -= __shfl_sync(MPP_FULL, incl - cnt, l);
        } else {
            target -= tot;
        }
    }
    if (row < 0) return MPP_NO_OBJECT;
    for (int b = 0; b < c.ny; b += 32) {
        const uint32_t msk = (b + lane < c.ny) ? __ldcg(c.mask + row * c.ny + b + lane) : 0u;
        const int cnt = __popc(msk);
        const int incl = warp_incl_scan(cnt, lane);
        const int tot = __shfl_sync(MPP_FULL, incl, 31);
        if (target < tot) {
            const uint32_t bal = __ballot_sync(MPP_FULL, incl > target);
            const int l = __ffs(bal) - 1;
            const int k = target - __shfl_sync(MPP_FULL, incl - cnt, l);
            const uint32_t m = __shfl_sync(MPP_FULL, msk, l);
            return (uint32_t)(row * c.ny + b + l) * 32u + __fns(m, 0, k + 1);
        }
        target -= tot;
    }
    return MPP_NO_OBJECT;
}

// One proposal of the reference's global kernel `kernel` against the current state.  Three Philox blocks r0..r2.
// Returns false when the kernel cannot act (empty configuration for kernels 1, 3..7): the empty perturbation.
template <typename R>
struct Drawn {
    bool has_rem, has_add;
    uint32_t rem_handle;
    Rec<R> rem, add;
    double d0, d1;
    int param_id, new_class;
};

template <typename R>
__device__ void draw_global(const Ctx<R> &c, const int *row_count, int n, int kernel, uint4 r0, uint4 r1, uint4 r2, int lane,
                            Drawn<R> *o) {
    o->has_rem = false; o->has_add = false; o->rem_handle = MPP_NO_OBJECT;
    o->d0 = 0; o->d1 = 0; o->param_id = 0; o->new_class = 0;
    if (kernel == 0) {
        o->add = uniform_birth(c, 0, c.H, 0, c.W, u01f(r0.z), u01f(r0.w), u01(r1.x, r1.y), u01(r1.z, r1.w), u01(r2.x, r2.y));
        o->has_add = true;
        return;
    }
    if (kernel == 2) {
        int x, y;
        sample_birth_pixel(c, u01(r0.z, r0.w), u01f(r1.x), u01f(r1.y), lane, &x, &y);
        o->add = data_birth_at(c, x, y, u01f(r1.z), u01f(r1.w), u01f(r2.x), lane);
        o->has_add = true;
        return;
    }
    if (n <= 0) return;
    o->rem_handle = pick_global(c, row_count, n, u01(r0.z, r0.w), lane);
    if (o->rem_handle == MPP_NO_OBJECT) { if (lane == 0) atomicOr(c.err, ERRF_NOT_FOUND); return; }
    o->rem = load_rec(c.recs + o->rem_handle);
    o->has_rem = true;
    if (kernel == 1 || kernel == 3) return;
    propose_move(c, kernel, o->rem, r1, u01f(r2.x), lane, &o->add, &o->d0, &o->d1, &o->param_id, &o->new_class);
    fill_unit_energies(c, o->add);
    o->has_add = true;
}
