// Proposal kernels of the RJMCMC sampler on device: forward / backward proposal
// probabilities (used by the replay and by the parallel sweep) and the device-side samplers (prefix-sum /
// inverse-CDF birth sampler, local-window translation sampler, 32-bin categorical mark sampler).
#pragma once
#include "mpp_device.cuh"

#define MPP_EPS 1e-16  // models/mpp/rjmcmc_sampler/rjmcmc.py:15

// float32 sum of a 32-element row in numpy's pairwise order (8 unrolled accumulators, then a 3-level tree):
// np.sum(density_map, axis=-1) in DataDrivenShapeTransformKernel.__init__ (transform_kernels.py:172)
__device__ __forceinline__ float row_sum_numpy(float v, int lane) {
    const int j = lane & 7;
    const float x1 = __shfl_sync(MPP_FULL, v, j + 8), x2 = __shfl_sync(MPP_FULL, v, j + 16), x3 = __shfl_sync(MPP_FULL, v, j + 24);
    const float a0 = __shfl_sync(MPP_FULL, v, j);
    float r = __fadd_rn(__fadd_rn(__fadd_rn(a0, x1), x2), x3);
    r = __fadd_rn(r, __shfl_xor_sync(MPP_FULL, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(MPP_FULL, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(MPP_FULL, r, 4));
    return __shfl_sync(MPP_FULL, r, 0);
}

// renormalised mark probability P_i[x,y,cls] / sum_k P_i[x,y,k] (float32), warp-cooperative: one coalesced 128-byte row
template <typename R>
__device__ __forceinline__ float mark_prob_n(const Ctx<R> &c, int i, int x, int y, int cls, int lane) {
    const float v = __ldg(mark_row(c, i, x, y) + lane);
    const float s = row_sum_numpy(v, lane);
    return __shfl_sync(MPP_FULL, __fdiv_rn(v, s), cls);
}

// RectangleSampler.get_point_density (shape_samplers.py:103-108)
template <typename R>
__device__ __forceinline__ double data_density(const Ctx<R> &c, int x, int y, uint32_t cls, int lane) {
    const float p0 = mark_prob_n(c, 0, x, y, cls_of(cls, 0), lane);
    const float p1 = mark_prob_n(c, 1, x, y, cls_of(cls, 1), lane);
    const float p2 = mark_prob_n(c, 2, x, y, cls_of(cls, 2), lane);
    const float detn = __fdiv_rn(__ldg(c.det + (size_t)x * c.W + y), c.det_sum);
    const float d32 = __fmul_rn(detn, __fmul_rn(__fmul_rn(p0, p1), p2));
    return (double)d32 * ((double)c.H * (double)c.W * 32768.0);
}

// DataDrivenTranslationKernel._move_density without p_kernel / n (transform_kernels.py:70-75,94-99)
template <typename R>
__device__ __forceinline__ double move_density(const Ctx<R> &c, int sx, int sy, int ex, int ey, int lane) {
    const int md = c.k.trl_max_delta;
    const int x0 = max(0, sx - md), x1 = min(sx + md + 1, c.H), y0 = max(0, sy - md), y1 = min(sy + md + 1, c.W);
    if (ex < x0 || ex >= x1 || ey < y0 || ey >= y1) return 0.0;
    const int wy = y1 - y0, tot = (x1 - x0) * wy;
    float acc = 0.f;
    for (int i = lane; i < tot; i += 32) {
        const int px = x0 + i / wy, py = y0 + i % wy;
        acc = __fadd_rn(acc, __fdiv_rn(__ldg(c.det + (size_t)px * c.W + py), c.det_sum));
    }
    acc = warp_sum(acc);
    const float e = __fdiv_rn(__ldg(c.det + (size_t)ex * c.W + ey), c.det_sum);
    return (double)__fdiv_rn(e, acc);
}

__device__ __forceinline__ double norm_pdf(double x, double sigma) {  // scipy.stats.norm.pdf(x, scale=sigma)
    const double z = x / sigma;
    return exp(-z * z / 2.0) / (sqrt(2.0 * 3.14159265358979323846) * sigma);
}

// Forward / backward probability of a proposal (R16-R19).  n = number of objects the kernel chose among
// (len(x) in the reference; the cell population in the parallel sweep), lambda = birth intensity
// (global_intensity, base_kernels.py:49; Lambda * Q_cell in the sweep), p = kernel choice probability.
template <typename R>
__device__ void proposal_probs(const Ctx<R> &c, int kernel, bool has_rem, const Rec<R> &rem, bool has_add,
                               const Rec<R> &add, double d0, double d1, int param_id, int new_class, double n,
                               double lambda, int lane, double *fwd, double *bwd) {
    const double p = c.k.p[kernel];
    switch (kernel) {
    case 0:
    case 2: {  // BirthKernel base_kernels.py:55-64
        const double dens = (kernel == 0 || !has_add) ? 1.0 : data_density(c, add.x, add.y, add.cls, lane);
        *fwd = (p * dens) / lambda;
        *bwd = p / (n + 1.0);
        return;
    }
    case 1:
    case 3: {  // DeathKernel base_kernels.py:94-115
        if (!has_rem) { *fwd = p; *bwd = p; return; }
        const double dens = kernel == 1 ? 1.0 : data_density(c, rem.x, rem.y, rem.cls, lane);
        *fwd = p / n;
        *bwd = (p * dens) / lambda;
        return;
    }
    default: break;
    }
    if (!has_rem) { *fwd = p; *bwd = p; return; }  // empty configuration: transform_kernels.py:49,58,108,...
    switch (kernel) {
    case 4: {  // GaussianTranslationKernel :42-58
        const double s = c.k.trl_sigma;
        *fwd = p * (norm_pdf(d0, s) * norm_pdf(d1, s)) / n;
        *bwd = p * (norm_pdf(-d0, s) * norm_pdf(-d1, s)) / n;
        return;
    }
    case 5: {  // DataDrivenTranslationKernel :101-116
        *fwd = p * move_density(c, rem.x, rem.y, add.x, add.y, lane) / n;
        *bwd = p * move_density(c, add.x, add.y, rem.x, rem.y, lane) / n;
        return;
    }
    case 6: {  // GaussianShapeTransformKernel :146-159
        const double f = p * norm_pdf(d0, c.k.trf_sigma[param_id]) / n;
        *fwd = f; *bwd = f;
        return;
    }
    default: {  // DataDrivenShapeTransformKernel :205-225
        const float v = __ldg(mark_row(c, param_id, rem.x, rem.y) + lane);
        const float s = row_sum_numpy(v, lane);
        const float pn = __fdiv_rn(v, s);
        *fwd = p * (double)__shfl_sync(MPP_FULL, pn, new_class) / n;
        *bwd = p * (double)__shfl_sync(MPP_FULL, pn, cls_of(rem.cls, param_id)) / n;
        return;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// K5 samplers.  All are warp-cooperative and return warp-uniform results.

// inverse CDF over 32 lane-held non-negative weights: first lane whose inclusive prefix exceeds u * total
// (numpy Generator.choice: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, side='right'))
__device__ __forceinline__ int warp_pick(float w, float u, int lane, float *total_out) {
    const float incl = warp_incl_scan(w, lane);
    const float total = __shfl_sync(MPP_FULL, incl, 31);
    const float t = u * total;
    const uint32_t b = __ballot_sync(MPP_FULL, incl > t && w > 0.f);
    if (total_out) *total_out = total;
    return b ? __ffs(b) - 1 : 31 - __clz(__ballot_sync(MPP_FULL, w > 0.f) | 1u);
}

// draws a pixel of the window [x0,x1) x [y0,y1) (at most 32 x 32) with probability proportional to det
template <typename R>
__device__ __forceinline__ void sample_window(const Ctx<R> &c, int x0, int x1, int y0, int y1, float u_row, float u_col,
                                              int lane, int *ox, int *oy, float *mass) {
    const int wx = x1 - x0, wy = y1 - y0;
    float rs = 0.f;
    if (lane < wx) {
        const float *row = c.det + (size_t)(x0 + lane) * c.W + y0;
        for (int j = 0; j < wy; ++j) rs += __ldg(row + j);
    }
    float total;
    const int r = warp_pick(rs, u_row, lane, &total);
    const float v = lane < wy ? __ldg(c.det + (size_t)(x0 + r) * c.W + y0 + lane) : 0.f;
    const int col = warp_pick(v, u_col, lane, nullptr);
    *ox = x0 + r; *oy = y0 + col;
    if (mass) *mass = total;
}

// categorical draw over the 32 classes of mark i at pixel (x,y) (shape_samplers.py:113-117)
template <typename R>
__device__ __forceinline__ int sample_mark_class(const Ctx<R> &c, int i, int x, int y, float u, int lane) {
    const float v = __ldg(mark_row(c, i, x, y) + lane);
    return warp_pick(v, u, lane, nullptr);
}

// global data-driven birth position: cell by binary search in the per-cell CDF, then the window sampler
// (replaces rng.choice over H*W pixels, utils/sampler2d.py:43-46, O(H*W) per draw in the reference)
template <typename R>
__device__ __forceinline__ void sample_birth_pixel(const Ctx<R> &c, double u_cell, float u_row, float u_col, int lane,
                                                   int *ox, int *oy) {
    const double t = u_cell * c.cell_cdf[c.ncell - 1];
    int lo = 0, hi = c.ncell - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c.cell_cdf[mid] > t) hi = mid; else lo = mid + 1;
    }
    const int ci = lo / c.ny, cj = lo % c.ny;
    sample_window(c, ci * 32, min(ci * 32 + 32, c.H), cj * 32, min(cj * 32 + 32, c.W), u_row, u_col, lane, ox, oy, nullptr);
}
