// R24: the optional split / merge kernels (rjmcmc_sampler/kernels/split_and_merge_kernels.py:14-178; make_kernels.py:145-161,
// off in both shipped configurations): proposal draws and forward / backward densities on the device, one warp per
// proposal.  The two-object Delta-energy telescopes into single-object Delta-energies (mpp_delta_batch).
#pragma once
#include "mpp_chain.cuh"

// ValueMapping.clip (mappings.py:39-43): cyclic marks wrap, the others saturate
__device__ __forceinline__ double sm_clip_mark(int i, double v) {
    const double vmax = mark_vmax(i);
    if (i == 2) { v = v - floor(v / vmax) * vmax; return (v < vmax && v >= 0.0) ? v : 0.0; }
    return fmin(fmax(v, 0.0), vmax);
}

// number of stored objects in the cells within one cell offset of (x, y): len(get_potential_neighbors(u, radius <= 32)) for a
// point that is not in the set (point_set.py:111-145); `euclid2` >= 0 adds the distance filter of get_neighbors (:147-149) and
// `exclude` drops one handle; when `out` is given the matching handles are listed in (x, y, uid) order is NOT applied here
template <typename R>
__device__ int sm_count_neighbors(const Ctx<R> &c, int x, int y, long long euclid2, uint32_t exclude, int lane, uint32_t *out, int cap) {
    const int iu = x >> 5, ju = y >> 5;
    int total = 0;
    for (int i = max(iu - 1, 0); i <= min(iu + 1, c.nx - 1); ++i)
        for (int j = max(ju - 1, 0); j <= min(ju + 1, c.ny - 1); ++j) {
            const int cell = j + i * c.ny;
            const uint32_t msk = __ldcg(c.mask + cell);
            bool in = (msk >> lane) & 1u;
            const uint32_t h = (uint32_t)cell * 32u + lane;
            if (in && h == exclude) in = false;
            if (in && euclid2 >= 0) {
                const int4 head = __ldcg(reinterpret_cast<const int4 *>(c.recs + h));
                const long long dx = head.x - x, dy = head.y - y;
                in = dx * dx + dy * dy <= euclid2;
            }
            const uint32_t b = __ballot_sync(MPP_FULL, in);
            if (in && out) { const int pos = total + __popc(b & ((1u << lane) - 1)); if (pos < cap) out[pos] = h; }
            total += __popc(b);
        }
    return total;
}

// SplitSampler.pdf (split_and_merge_kernels.py:33-36): uniform on the disc of radius r times three normal densities
__device__ __forceinline__ double sm_split_pdf(double radius, const double *sig, const double *shape_delta) {
    double p = 1.0 / (3.14159265358979323846 * radius * radius);
    for (int i = 0; i < 3; ++i) {
        const double z = shape_delta[i] / sig[i];
        p *= exp(-0.5 * z * z) / (2.5066282746310002 * sig[i]);
    }
    return p;
}

template <typename R>
__global__ void k_sample_split_merge(Ctx<R> c, const int *__restrict__ row_count, int kind, double radius, double s0, double s1, double s2,
                                     uint64_t seed, uint64_t offset, mpp_split_merge *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    __shared__ uint32_t nb[96];
    Philox rng(seed, (uint32_t)offset, (uint32_t)(offset >> 32), 0x5b17u);
    const uint4 r0 = rng.next(), r1 = rng.next();
    mpp_split_merge o;
    memset(&o, 0, sizeof(o));
    o.kind = kind; o.rem_uid[0] = o.rem_uid[1] = MPP_NO_OBJECT; o.n_add = 0; o.n_neighbors = -1;
    const int n = __ldcg(c.n_objects);
    const double sig[3] = {s0 * 32.0, s1 * 1.0, s2 * 3.14159265358979323846};  // sigma * mapping.range (split_and_merge_kernels.py:21)
    if ((kind == 8 && n > 0) || (kind == 9 && n > 1)) {
        const uint32_t h0 = pick_global(c, row_count, n, u01(r0.x, r0.y), lane);   // PointsSet.random_choice point_set.py:176-185
        const Rec<R> p0 = load_rec(c.recs + h0);
        o.rem_uid[0] = p0.uid; o.rem_x[0] = p0.x; o.rem_y[0] = p0.y;
        if (kind == 8) {
            // position delta: uniform on [0, r)^2, redrawn while outside the disc (the reference's rejection loop, :27-29);
            // shape deltas: three independent normals
            double dx = 0, dy = 0;
            for (int t = 0; t < 64; ++t) {
                const uint4 q = rng.next();
                dx = u01(q.x, q.y) * radius; dy = u01(q.z, q.w) * radius;
                if (dx * dx + dy * dy <= radius * radius) break;
            }
            const uint4 g0 = rng.next(), g1 = rng.next();
            const double ua[3] = {u01(g0.x, g0.y), u01(g0.z, g0.w), u01(g1.x, g1.y)}, ub[3] = {u01(r1.x, r1.y), u01(r1.z, r1.w), u01(g1.z, g1.w)};
            o.pos_delta[0] = dx; o.pos_delta[1] = dy;
            const double marks[3] = {(double)p0.size, (double)p0.ratio, (double)p0.angle};
            double a0[3], a1[3];
            for (int i = 0; i < 3; ++i) {
                const double d = sqrt(-2.0 * log(ua[i])) * cospi(2.0 * ub[i]) * sig[i];
                o.shape_delta[i] = d;
                a0[i] = sm_clip_mark(i, marks[i] - d); a1[i] = sm_clip_mark(i, marks[i] + d);
            }
            o.n_add = 2;
            o.add_x[0] = min(max((int)((double)p0.x - dx), 0), c.H - 1); o.add_y[0] = min(max((int)((double)p0.y - dy), 0), c.W - 1);
            o.add_x[1] = min(max((int)((double)p0.x + dx), 0), c.H - 1); o.add_y[1] = min(max((int)((double)p0.y + dy), 0), c.W - 1);
            o.add_size[0] = a0[0]; o.add_ratio[0] = a0[1]; o.add_angle[0] = a0[2];
            o.add_size[1] = a1[0]; o.add_ratio[1] = a1[1]; o.add_angle[1] = a1[2];
        } else {
            // neighbours of p0 within the radius (get_neighbors, Euclidean), listed in (x, y, uid) order; one is drawn uniformly
            const long long r2 = (long long)floor(radius * radius);
            const int cnt = sm_count_neighbors(c, p0.x, p0.y, r2, h0, lane, nb, 96);
            __syncwarp();
            o.n_neighbors = cnt;
            if (cnt > 0 && cnt <= 96) {
                const int want = min(cnt - 1, (int)(u01(r0.z, r0.w) * (double)cnt));
                // rank of every listed neighbour in (x, y, uid) order
                uint32_t pickh = MPP_NO_OBJECT;
                for (int b = 0; b < cnt; b += 32) {
                    const int k = b + lane;
                    bool hit = false;
                    if (k < cnt) {
                        const int4 hk = __ldcg(reinterpret_cast<const int4 *>(c.recs + nb[k]));
                        int rank = 0;
                        for (int v = 0; v < cnt; ++v) {
                            const int4 hv = __ldcg(reinterpret_cast<const int4 *>(c.recs + nb[v]));
                            rank += (hv.x < hk.x) || (hv.x == hk.x && (hv.y < hk.y || (hv.y == hk.y && (uint32_t)hv.w < (uint32_t)hk.w)));
                        }
                        hit = rank == want;
                    }
                    const uint32_t bal = __ballot_sync(MPP_FULL, hit);
                    if (bal) pickh = nb[b + __ffs(bal) - 1];
                }
                const Rec<R> p1 = load_rec(c.recs + pickh);
                o.rem_uid[1] = p1.uid; o.rem_x[1] = p1.x; o.rem_y[1] = p1.y;
                o.n_add = 1;
                // the merged object is the average of the two; the reference clips BOTH coordinates with shape[0] (:142-143)
                o.add_x[0] = min(max((int)(((double)p0.x + (double)p1.x) / 2.0), 0), c.H - 1);
                o.add_y[0] = min(max((int)(((double)p0.y + (double)p1.y) / 2.0), 0), c.H - 1);
                o.add_size[0] = sm_clip_mark(0, ((double)p0.size + (double)p1.size) / 2.0);
                o.add_ratio[0] = sm_clip_mark(1, ((double)p0.ratio + (double)p1.ratio) / 2.0);
                o.add_angle[0] = sm_clip_mark(2, ((double)p0.angle + (double)p1.angle) / 2.0);
                o.pos_delta[0] = ((double)p0.x - (double)p1.x) / 2.0; o.pos_delta[1] = ((double)p0.y - (double)p1.y) / 2.0;
                o.shape_delta[0] = ((double)p0.size - (double)p1.size) / 2.0; o.shape_delta[1] = ((double)p0.ratio - (double)p1.ratio) / 2.0;
                o.shape_delta[2] = ((double)p0.angle - (double)p1.angle) / 2.0;
            }
        }
    }
    o.u = u01(r1.x ^ 0x9e3779b9u, r1.w);
    if (lane == 0) *out = o;
}

// forward / backward densities of a split (kind 8) or merge (kind 9) perturbation against the current state
// (split_and_merge_kernels.py:76-107, 151-178).  out[0] = forward, out[1] = backward.
template <typename R>
__global__ void k_split_merge_probs(Ctx<R> c, const mpp_split_merge *__restrict__ in, double p_split, double p_merge, double radius,
                                    double s0, double s1, double s2, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const mpp_split_merge p = *in;
    const int n = __ldcg(c.n_objects);
    const double sig[3] = {s0 * 32.0, s1 * 1.0, s2 * 3.14159265358979323846};
    double fwd, bwd;
    if (p.kind == 8) {
        fwd = n != 0 ? p_split * ((1.0 / (double)n) * sm_split_pdf(radius, sig, p.shape_delta)) / c.k.intensity : p_split;
        const int n1 = n + 1;
        if (n1 > 1) {
            const int nb0 = sm_count_neighbors(c, p.add_x[0], p.add_y[0], -1, MPP_NO_OBJECT, lane, (uint32_t *)nullptr, 0) + 1;
            const int nb1 = sm_count_neighbors(c, p.add_x[1], p.add_y[1], -1, MPP_NO_OBJECT, lane, (uint32_t *)nullptr, 0) + 1;
            bwd = p_merge * ((1.0 / (double)n1) * (1.0 / (double)nb0) + (1.0 / (double)n1) * (1.0 / (double)nb1));
        } else bwd = p_merge;
    } else {
        fwd = (n > 1 && p.n_neighbors != 0) ? p_merge * ((1.0 / (double)n) * (1.0 / (double)p.n_neighbors)) : p_merge;
        const int n1 = n - 1;
        if (n1 != 0) bwd = p.rem_uid[1] == MPP_NO_OBJECT ? p_split : p_split * ((1.0 / (double)n1) * sm_split_pdf(radius, sig, p.shape_delta)) / c.k.intensity;
        else bwd = p_split;
    }
    if (lane == 0) { out[0] = fwd; out[1] = bwd; }
}
