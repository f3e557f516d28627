// K6 v2: parallel colour-sweep sampler with shared-memory resident windows.
//
// One CTA owns one 32x32-px sampling window for the whole visit.  The objects within 64 px of the window are staged
// ONCE in shared memory (geometry, unit energies, and for every object that a move in the window can affect its
// top-2 partner reductions), then `per_visit` Metropolis-Hastings-Green proposals run from shared memory; only the
// map gathers of a proposed object touch L2/HBM.  The NW warps of the CTA evaluate NW consecutive proposals of the
// chain speculatively against the same state; the first accepted one is committed and the later ones are discarded and
// re-drawn from their own Philox counters, so the chain is bit-identical for every NW ("speculative moves").
//
// Windows are the cells of the 32-px grid shifted by a per-sweep random offset (ox, oy), coloured 3x3: windows of one
// colour are >= 65 px apart (> the 64-px reach of a Delta-energy), so they commute.  Moves that leave the window are
// rejected (symmetric restriction); the random shift lets objects cross every boundary over time.
// Reference rows: R15-R21 (rjmcmc.py:83-164, kernels/*.py); energies as in mpp_device.cuh.
#pragma once
#include "mpp_chain.cuh"

#define W2_K 128         // staged objects per window (window +- 64 px)
#define W2_MAXW 8        // warps per window (speculation depth)
#define W2_PRE 128       // proposals per visit that can be drawn ahead (= the largest proposals_per_visit)
#define W2_EPS 1e-16f
// quotients of the Green ratio (proposal densities, Delta E / T): the approximate division (2 ulp) is far inside the float32
// noise of log alpha and saves the refinement steps + range check of the IEEE one on the critical path of every proposal
#ifndef MPP_EXACT_DIV
#define W2_DIV(a, b) __fdividef((a), (b))
#else
#define W2_DIV(a, b) ((a) / (b))
#endif
#define W2_SCRATCH (32 + 2 * W2_K)  // per-warp scratch (elements): window row masses of the pre-draw + pair-value stash

enum : unsigned char { W2_ALIVE = 1, W2_WIN = 2, W2_INNER = 4 };

template <typename R>
struct WinState {
    int n;        // staged entries (alive or free)
    int n_win;    // alive objects inside the window
    int x0, x1, y0, y1;   // window pixel box (clipped to the image)
    int cx0, cy0;         // first storage cell (row, col) of the 2x2 block the window overlaps
    uint32_t cmask[4];    // private copy of the occupancy masks of those storage cells (NO cell: 0xffffffff)
    int ccell[4];         // their linear indices (-1: outside the grid)
    uint32_t uid_base;
    int dn, n_acc, n_birth, n_death, n_eval, n_done;
    float row_mass[32];   // detection mass of each window row;  win_mass = their sum
    double win_mass;
    // per-visit constants
    float pkf[8];         // kernel-choice probabilities (make_kernels.py:76-86)
    float pk_e0, pk_e2;   // ... of the two birth kernels when the window is empty (births only)
    float lam_unif, lam_data;  // birth intensity of the window: Lambda * (share of the uniform / data-driven proposal mass)
    float dens_scale;     // H * W * 32^3 / sum(det): RectangleSampler.get_point_density (shape_samplers.py:88,103-108)
    int masks_dirty;
    int x[W2_K], y[W2_K];
    uint32_t cls[W2_K], handle[W2_K], uid[W2_K];
    R size[W2_K], ratio[W2_K], angle[W2_K];
    R hl[W2_K], hw[W2_K], ca[W2_K], sa[W2_K], rad[W2_K];
    R pos[W2_K], tm0[W2_K], tm1[W2_K], tm2[W2_K];      // unit data terms as the combinator sees them
    R dm0[W2_K], dm1[W2_K], dm2[W2_K];                  // per-mark energies (window objects; legacy: before the mean)
    float detv[W2_K], pn0[W2_K], pn1[W2_K], pn2[W2_K];  // det value and normalised mark probabilities (window objects)
    R ov1[W2_K], ov2[W2_K], al1[W2_K], al2[W2_K];
    R fa[W2_K], fg[W2_K];  // the combinator's gated linear form per entry: f = fa + fg * (c_ov * ov + c_al * al) (+ logistic)
    R fcur[W2_K];          // current combined energy f(unit terms, ov1, al1) of every entry whose reductions are maintained (W2_INNER):
                           // the "before" half of every Delta-energy, kept up to date by the commits instead of being recomputed
    short aov[W2_K], aal[W2_K], aov2[W2_K], aal2[W2_K];  // staged indices of the best / second-best partners
    unsigned char flags[W2_K];
    uint32_t order[W2_K];  // staging scratch: handles in canonical order
    short winlist[W2_K];   // SIMT mode: staged indices of the alive window objects
    // speculation results, one slot per warp
    int res_accept[2][W2_MAXW], res_eval[W2_MAXW];  // res_accept: double-buffered by round parity (a round without a commit has one barrier)
    // per-visit statistics (mpp_window_stats): [0..15] evaluated by (window empty ? 0 : 8) + kernel, [16..31] accepted likewise,
    // [32] identity proposals accepted, [33] visits, [34] visits that found the window empty
    int kstat[MPP_WINDOW_STATS];
    // proposals drawn ahead (warp mode): random words of every proposal of the visit, its kernel under the two mixtures
    // (index 0: empty window, births only; 1: the reference mixture) and, where that kernel is a birth, the candidate
    uint32_t pq[8][W2_PRE];
    unsigned char pkern[2][W2_PRE];
    int pc_x[2][W2_PRE], pc_y[2][W2_PRE];
    uint32_t pc_cls[2][W2_PRE];
    R pc_size[2][W2_PRE], pc_ratio[2][W2_PRE], pc_angle[2][W2_PRE], pc_hl[2][W2_PRE], pc_hw[2][W2_PRE], pc_ca[2][W2_PRE], pc_sa[2][W2_PRE];
    R pc_pos[2][W2_PRE], pc_dm0[2][W2_PRE], pc_dm1[2][W2_PRE], pc_dm2[2][W2_PRE];
    float pc_detv[2][W2_PRE], pc_pn0[2][W2_PRE], pc_pn1[2][W2_PRE], pc_pn2[2][W2_PRE];
};

template <typename R>
struct Cand {  // the object a proposal wants to add (warp-uniform registers)
    int x, y;
    uint32_t cls;
    R size, ratio, angle, hl, hw, ca, sa;
    R pos, dm0, dm1, dm2;
    float detv, pn0, pn1, pn2;
};

template <typename R>
__device__ __forceinline__ void shape_terms(const ModelDev &m, R dm0, R dm1, R dm2, R *t0, R *t1, R *t2) {
    if (m.setup == MPP_SETUP_LEGACY) {  // float(np.mean([d0,d1,d2])) data_energies.py:43
        // sum / 3, correctly rounded like the IEEE quotient (reciprocal product + one FMA correction step: exact for the
        // divisor 3) without the range check and slow path of the division
        const float sum3 = __fadd_rn(__fadd_rn((float)dm0, (float)dm1), (float)dm2);
        const float q3 = __fmul_rn(sum3, 0.333333343267440796f);
        *t0 = (R)__fmaf_rn(__fmaf_rn(-3.0f, q3, sum3), 0.333333343267440796f, q3);
        *t1 = 0; *t2 = 0;
    } else {
        *t0 = dm0; *t1 = dm1; *t2 = dm2;
    }
}

// per-mark energy of the sampler: the legacy remap -2*sigmoid(c*p + b) + 1 == -tanh((c*p + b) / 2) with the fast exponential
// (absolute error < 2e-7, far inside the 1e-5 parity budget; the exact float32 sequence of the reference is kept in
// legacy_remap_f32 for the parity entry points)
__device__ __forceinline__ float mark_energy_f32(const ModelDev &m, int i, float p) {
    if (m.premapped) return p;
    if (m.setup != MPP_SETUP_LEGACY) return -p;
    const float z = fmaf(p, m.coef[i], m.icpt[i]);
    return 1.0f - __fdividef(2.0f, 1.0f + __expf(-z));
}

// det value, per-mark energies and normalised mark probabilities of classes `cls` at pixel (x, y): seven independent gathers
template <typename R>
__device__ __forceinline__ void gather_pixel(const Ctx<R> &c, int x, int y, uint32_t cls, float *detv, float *pn, float *dm) {
    const size_t pix = (size_t)x * c.W + y, plane = (size_t)c.H * c.W;
    const float d = __ldg(c.det + pix);
    const float p0 = __ldg(mark_row(c, 0, x, y) + cls_of(cls, 0)), p1 = __ldg(mark_row(c, 1, x, y) + cls_of(cls, 1)),
                p2 = __ldg(mark_row(c, 2, x, y) + cls_of(cls, 2));
    const float s0 = __ldg(c.marksum + pix), s1 = __ldg(c.marksum + plane + pix), s2 = __ldg(c.marksum + 2 * plane + pix);
    *detv = d;
    pn[0] = __fdividef(p0, s0); pn[1] = __fdividef(p1, s1); pn[2] = __fdividef(p2, s2);
    dm[0] = mark_energy_f32(c.m, 0, p0); dm[1] = mark_energy_f32(c.m, 1, p1); dm[2] = mark_energy_f32(c.m, 2, p2);
}

// out-of-line copy for the warp-uniform callers of the rounds (every lane issues the same seven gathers: one transaction each)
// (the seven values come back by value, in registers: through pointers they forced the caller's candidate onto the local stack)
struct PixelInfo { float detv, pn0, pn1, pn2, dm0, dm1, dm2; };
template <typename R>
__device__ __noinline__ PixelInfo pixel_info_v(const Ctx<R> &c, int x, int y, uint32_t cls) {
    PixelInfo o;
    float pn[3], dm[3];
    gather_pixel(c, x, y, cls, &o.detv, pn, dm);
    o.pn0 = pn[0]; o.pn1 = pn[1]; o.pn2 = pn[2]; o.dm0 = dm[0]; o.dm1 = dm[1]; o.dm2 = dm[2];
    return o;
}
template <typename R>
__device__ __forceinline__ void pixel_info(const Ctx<R> &c, int x, int y, uint32_t cls, int lane, float *detv, float *pn, float *dm) {
    (void)lane;
    const PixelInfo o = pixel_info_v<R>(c, x, y, cls);
    *detv = o.detv; pn[0] = o.pn0; pn[1] = o.pn1; pn[2] = o.pn2; dm[0] = o.dm0; dm[1] = o.dm1; dm[2] = o.dm2;
}

template <typename R>
__device__ __forceinline__ Geo<R> geo_w(const WinState<R> &w, int k) {
    Geo<R> g; g.x = w.x[k]; g.y = w.y[k]; g.hl = w.hl[k]; g.hw = w.hw[k]; g.ca = w.ca[k]; g.sa = w.sa[k];
    return g;
}

// RectangleSampler.get_point_density (shape_samplers.py:103-108) from staged factors
template <typename R>
__device__ __forceinline__ float dens_of(const WinState<R> &w, float detv, float pn0, float pn1, float pn2) {
    return detv * (pn0 * pn1 * pn2) * w.dens_scale;
}

// overlap-kind pair value between staged entry k and an object with geometry gb / bounding radius rad_b; the
// bounding-circle test is done inline so that the (out-of-line) polygon clip only runs for pairs that can intersect
template <typename R>
__device__ __forceinline__ R pair_ov_w(const ModelDev &m, const WinState<R> &w, int k, const Geo<R> &gb, R rad_b, int d2, R *sx, R *sy) {
    if (m.setup == MPP_SETUP_TOY) return d2 <= m.toy_d2 ? (R)m.toy_pair : (R)0;
    const R rr = w.rad[k] + rad_b;
    if ((R)d2 > rr * rr * (R)1.0001) return (R)0;
    return overlap_energy_v<R, true>(w.x[k], w.y[k], w.hl[k], w.hw[k], w.ca[k], w.sa[k], gb.x, gb.y, gb.hl, gb.hw, gb.ca, gb.sa);
}

// top-2 partner reductions of staged entry k over every other alive staged entry (one lane, serial loop)
template <typename R>
__device__ __noinline__ void recompute_top2(const ModelDev &m, WinState<R> &w, int k, bool do_ov, bool do_al, R *sx, R *sy) {
    R o1 = 0, o2 = 0, a1 = 0, a2 = 0;
    int ao = -1, aa = -1, ao2 = -1, aa2 = -1;
    const Geo<R> gk = geo_w(w, k);
    const R rk = w.rad[k];
    const int n = w.n;
    for (int v = 0; v < n; ++v) {
        if (v == k || !(w.flags[v] & W2_ALIVE)) continue;
        const int dx = w.x[v] - gk.x, dy = w.y[v] - gk.y, d2 = dx * dx + dy * dy;
        if (d2 > m.max_d2) continue;
        if (do_ov && d2 <= m.ov_d2) {
            const R o = pair_ov_w(m, w, v, gk, rk, d2, sx, sy);
            if (o > o1) { o2 = o1; ao2 = ao; o1 = o; ao = v; } else if (o > o2) { o2 = o; ao2 = v; }
        }
        if (do_al && d2 <= m.al_d2) {
            const R a = align_magnitude(gk, geo_w(w, v), m.rewarding);
            if (a > a1) { a2 = a1; aa2 = aa; a1 = a; aa = v; } else if (a > a2) { a2 = a; aa2 = v; }
        }
    }
    if (do_ov) { w.ov1[k] = o1; w.ov2[k] = o2; w.aov[k] = (short)ao; w.aov2[k] = (short)ao2; }
    if (do_al) { w.al1[k] = a1; w.al2[k] = a2; w.aal[k] = (short)aa; w.aal2[k] = (short)aa2; }
}

// the same reductions by a group of 8 lanes (staging): the partners are spread over the lanes, the per-lane top-2 are merged by
// three xor-shuffle steps.  `gmask`: the group's lanes; j: lane index within the group.  Result in lane j == 0.
template <typename R>
__device__ __forceinline__ void merge_top2(R &o1, int &i1, R &o2, int &i2, R b1, int j1, R b2, int j2) {
    if (b1 > o1) {
        if (o1 >= b2) { o2 = o1; i2 = i1; } else { o2 = b2; i2 = j2; }
        o1 = b1; i1 = j1;
    } else if (b1 > o2) { o2 = b1; i2 = j1; }
}
template <typename R>
__device__ __forceinline__ void group_top2(const ModelDev &m, WinState<R> &w, int k, int j, uint32_t gmask, R *sx, R *sy) {
    R o1 = 0, o2 = 0, a1 = 0, a2 = 0;
    int ao = -1, aa = -1, ao2 = -1, aa2 = -1;
    const Geo<R> gk = geo_w(w, k);
    const R rk = w.rad[k];
    const int n = w.n;
    for (int v = j; v < n; v += 8) {
        if (v == k || !(w.flags[v] & W2_ALIVE)) continue;
        const int dx = w.x[v] - gk.x, dy = w.y[v] - gk.y, d2 = dx * dx + dy * dy;
        if (d2 > m.max_d2) continue;
        if (d2 <= m.ov_d2) {
            const R o = pair_ov_w(m, w, v, gk, rk, d2, sx, sy);
            if (o > o1) { o2 = o1; ao2 = ao; o1 = o; ao = v; } else if (o > o2) { o2 = o; ao2 = v; }
        }
        if (d2 <= m.al_d2) {
            const R a = align_magnitude(gk, geo_w(w, v), m.rewarding);
            if (a > a1) { a2 = a1; aa2 = aa; a1 = a; aa = v; } else if (a > a2) { a2 = a; aa2 = v; }
        }
    }
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
        const R bo1 = __shfl_xor_sync(gmask, o1, off), bo2 = __shfl_xor_sync(gmask, o2, off);
        const int bi1 = __shfl_xor_sync(gmask, ao, off), bi2 = __shfl_xor_sync(gmask, ao2, off);
        const R ba1 = __shfl_xor_sync(gmask, a1, off), ba2 = __shfl_xor_sync(gmask, a2, off);
        const int bj1 = __shfl_xor_sync(gmask, aa, off), bj2 = __shfl_xor_sync(gmask, aa2, off);
        merge_top2(o1, ao, o2, ao2, bo1, bi1, bo2, bi2);
        merge_top2(a1, aa, a2, aa2, ba1, bj1, ba2, bj2);
    }
    if (j == 0) {
        w.ov1[k] = o1; w.ov2[k] = o2; w.aov[k] = (short)ao; w.aov2[k] = (short)ao2;
        w.al1[k] = a1; w.al2[k] = a2; w.aal[k] = (short)aa; w.aal2[k] = (short)aa2;
    }
}

#ifdef MPP_V_FOBJ
#define MPP_FOBJ_INL __noinline__
#else
#define MPP_FOBJ_INL __forceinline__
#endif
// (out of line: three inlined copies of the unrolled 32-step scan cost 2 % through instruction-cache misses)
#define MPP_SCAN_INL __noinline__
#ifdef MPP_V_DELTA
#define MPP_DELTA_INL __noinline__
#else
#define MPP_DELTA_INL
#endif
// unit part of the combinator's gated linear form (combine_fast) for staged entry k: everything but the two partner terms
template <typename R>
__device__ __forceinline__ void unit_form_vals(const ModelDev &m, R tm0, R tm1, R tm2, R hl, R hw, R ratio, R pos, R *fa, R *fg) {
    const R uin = (R)m.c_m0 * tm0 + (R)m.c_m1 * tm1 + (R)m.c_m2 * tm2 + (R)m.c_area * area_prior_fast<R>(m, hl, hw) +
                  (R)m.c_ratio * r_abs((R)m.f_target_ratio - ratio);
    const R g = (m.gate && !(pos <= (R)m.gate_thr)) ? (R)0 : (R)1;
    *fa = (R)m.c_pos * pos + g * uin + (R)m.c_0;
    *fg = g;
}
template <typename R>
__device__ __forceinline__ void unit_form(const ModelDev &m, WinState<R> &w, int k) {
    R fa, fg;
    unit_form_vals<R>(m, w.tm0[k], w.tm1[k], w.tm2[k], w.hl[k], w.hw[k], w.ratio[k], w.pos[k], &fa, &fg);
    w.fa[k] = fa; w.fg[k] = fg;
}
template <typename R>
__device__ MPP_FOBJ_INL R f_obj(const ModelDev &m, const WinState<R> &w, int k, R ov, R al) {
    const R e = w.fa[k] + w.fg[k] * ((R)m.c_ov * ov + (R)m.c_al * ((m.rewarding ? (R)-1 : (R)1) * al));
    return m.logistic ? r_logistic_pm1(e) : e;
}

// Delta-energy of removing staged entry r (r < 0: none) and/or adding `a` (has_add), from the staged state only.
// po / pa (per-warp scratch, W2_K entries each) receive the overlap / alignment pair values between every staged
// entry and `a` (0 where out of reach) so that an accepted proposal can update the top-2 reductions incrementally.
template <typename R>
__device__ MPP_DELTA_INL R delta_staged(const ModelDev &m, const WinState<R> &w, int r, bool has_add, const Cand<R> &a, int lane, R *sx, R *sy, R *po, R *pa) {
    Geo<R> ga; ga.x = a.x; ga.y = a.y; ga.hl = a.hl; ga.hw = a.hw; ga.ca = a.ca; ga.sa = a.sa;
    const R rad_a = has_add ? r_sqrt_fast(a.hl * a.hl + a.hw * a.hw) : (R)0;
    const int rx = r >= 0 ? w.x[r] : 0, ry = r >= 0 ? w.y[r] : 0;
    R acc = 0, ov_add = 0, al_add = 0;
    const int n = w.n;
#pragma unroll 1
    for (int k = lane; k < n; k += 32) {
        R o = 0, al = 0;
        if (k != r && (w.flags[k] & W2_ALIVE)) {
            bool touched = false;
            R ov_b = w.ov1[k], al_b = w.al1[k], ov_a = ov_b, al_a = al_b;
            if (r >= 0) {
                const int dx = w.x[k] - rx, dy = w.y[k] - ry;
                if (dx * dx + dy * dy <= m.max_d2) {
                    touched = true;
                    if (w.aov[k] == r) ov_a = w.ov2[k];
                    if (w.aal[k] == r) al_a = w.al2[k];
                }
            }
            if (has_add) {
                const int dx = w.x[k] - a.x, dy = w.y[k] - a.y, d2 = dx * dx + dy * dy;
                if (d2 <= m.max_d2) {
                    touched = true;
                    if (d2 <= m.ov_d2) { o = pair_ov_w(m, w, k, ga, rad_a, d2, sx, sy); ov_a = r_max(ov_a, o); ov_add = r_max(ov_add, o); }
                    if (d2 <= m.al_d2) { al = align_magnitude(geo_w(w, k), ga, m.rewarding); al_a = r_max(al_a, al); al_add = r_max(al_add, al); }
                }
            }
            if (touched) acc += f_obj(m, w, k, ov_a, al_a) - w.fcur[k];
        }
        if (has_add) { po[k] = o; pa[k] = al; }
    }
    acc = warp_sum_sparse(acc);
    if (has_add) {
        ov_add = warp_max_nonneg(ov_add); al_add = warp_max_nonneg(al_add);
        Terms<R> t;
        t.pos = a.pos;
        shape_terms<R>(m, a.dm0, a.dm1, a.dm2, &t.m0, &t.m1, &t.m2);
        t.ov = ov_add; t.al = (m.rewarding ? (R)-1 : (R)1) * al_add;
        t.area = area_prior_fast<R>(m, a.hl, a.hw);
        t.ratio = r_abs((R)m.f_target_ratio - a.ratio);
        acc += combine_fast(m, t);
    }
    if (r >= 0) acc -= w.fcur[r];
    return acc;
}

// brute-force version of delta_staged (debug): recomputes every reduction without the top-2 machinery
template <typename R>
__device__ R delta_brute(const ModelDev &m, const WinState<R> &w, int r, bool has_add, const Cand<R> &a, int lane, R *sx, R *sy) {
    Geo<R> ga; ga.x = a.x; ga.y = a.y; ga.hl = a.hl; ga.hw = a.hw; ga.ca = a.ca; ga.sa = a.sa;
    R acc = 0;
    const int n = w.n;
#pragma unroll 1
    for (int k = lane; k < n; k += 32) {
        if (!(w.flags[k] & W2_ALIVE)) continue;
        R ob = 0, ab = 0, oa = 0, aa = 0;
        const Geo<R> gk = geo_w(w, k);
        for (int v = 0; v < n; ++v) {
            if (v == k || !(w.flags[v] & W2_ALIVE)) continue;
            const int dx = w.x[v] - gk.x, dy = w.y[v] - gk.y, d2 = dx * dx + dy * dy;
            if (d2 > m.max_d2) continue;
            const Geo<R> gv = geo_w(w, v);
            R o = 0, al = 0;
            if (d2 <= m.ov_d2) o = pair_overlap(m, gk, gv, d2, sx, sy);
            if (d2 <= m.al_d2) al = align_magnitude(gk, gv, m.rewarding);
            ob = r_max(ob, o); ab = r_max(ab, al);
            if (v != r) { oa = r_max(oa, o); aa = r_max(aa, al); }
        }
        if (has_add && k != r) {
            const int dx = a.x - gk.x, dy = a.y - gk.y, d2 = dx * dx + dy * dy;
            if (d2 <= m.ov_d2) oa = r_max(oa, pair_overlap(m, gk, ga, d2, sx, sy));
            if (d2 <= m.al_d2) aa = r_max(aa, align_magnitude(gk, ga, m.rewarding));
        }
        // only objects that can be affected have meaningful unit terms; the others cancel exactly (same ov/al)
        if (k == r) acc -= f_obj(m, w, k, ob, ab);
        else if (oa != ob || aa != ab) acc += f_obj(m, w, k, oa, aa) - f_obj(m, w, k, ob, ab);
    }
    acc = warp_sum(acc);
    if (has_add) {
        R oa = 0, aa = 0;
        for (int v = lane; v < n; v += 32) {
            if (v == r || !(w.flags[v] & W2_ALIVE)) continue;
            const int dx = w.x[v] - a.x, dy = w.y[v] - a.y, d2 = dx * dx + dy * dy;
            if (d2 > m.max_d2) continue;
            const Geo<R> gv = geo_w(w, v);
            if (d2 <= m.ov_d2) oa = r_max(oa, pair_overlap(m, ga, gv, d2, sx, sy));
            if (d2 <= m.al_d2) aa = r_max(aa, align_magnitude(ga, gv, m.rewarding));
        }
        oa = warp_max(oa); aa = warp_max(aa);
        Terms<R> t;
        t.pos = a.pos;
        shape_terms<R>(m, a.dm0, a.dm1, a.dm2, &t.m0, &t.m1, &t.m2);
        t.ov = oa; t.al = (m.rewarding ? (R)-1 : (R)1) * aa;
        t.area = area_prior_fast<R>(m, a.hl, a.hw);
        t.ratio = r_abs((R)m.f_target_ratio - a.ratio);
        acc += combine_fast(m, t);
    }
    return acc;
}

// out-of-line copies of the two most replicated warp primitives of the sampler (instruction-cache footprint: with many
// windows resident per SM the kernel is instruction-fetch bound, not issue bound)
struct PickTotal { int idx; float total; };  // (by value, in registers: an out-pointer to an out-of-line function goes through the local stack)
__device__ __noinline__ PickTotal warp_pick_nv(float w, float u, int lane) {
    PickTotal o;
    o.idx = warp_pick(w, u, lane, &o.total);
    return o;
}
__device__ __forceinline__ int warp_pick_ni(float w, float u, int lane, float *total_out) {
    const PickTotal o = warp_pick_nv(w, u, lane);
    if (total_out) *total_out = o.total;
    return o.idx;
}
__device__ __noinline__ void philox2(uint64_t seed, uint32_t c1, uint32_t c2, uint32_t c3, uint4 *q0, uint4 *q1) {
    Philox rng(seed, c1, c2, c3);
    *q0 = rng.next(); *q1 = rng.next();
}

__device__ __forceinline__ void box_muller_f(uint32_t a, uint32_t b, float *n0, float *n1) {
    const float u1 = u01f(a), u2 = u01f(b);
    // hardware square root and sine / cosine on an argument reduced to [-pi, pi): absolute error < 1e-6 on a standard normal
    // draw, i.e. 1e-5 px / 1e-6 of a mark range after scaling -- the IEEE sqrtf + sincospif cost ~70 instructions and a slow-path call
    const float r = r_sqrt_fast(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(3.14159265358979f * (2.0f * u2 - (u2 >= 0.5f ? 2.0f : 0.0f)), &sn, &cs);
    *n0 = r * cs; *n1 = r * sn;
}

// kernel-choice probabilities of a window holding n objects: the reference mixture (make_kernels.py:76-86) when
// n >= 1; births only (renormalised) when the window is empty, where every other kernel is the empty perturbation.
template <typename R>
__device__ __forceinline__ float pk_of(const WinState<R> &w, int kernel, int n) {
    if (n > 0) return w.pkf[kernel];
    return kernel == 0 ? w.pk_e0 : (kernel == 2 ? w.pk_e2 : 0.f);
}

// j-th (0-based) alive window object among the staged entries
template <typename R>
__device__ __forceinline__ int pick_window_object(const WinState<R> &w, int j, int lane) {
    const int n = w.n;
#pragma unroll 1
    for (int b = 0; b < n; b += 32) {
        const int k = b + lane;
        const bool in = k < n && (w.flags[k] & (W2_ALIVE | W2_WIN)) == (W2_ALIVE | W2_WIN);
        const uint32_t bal = __ballot_sync(MPP_FULL, in);
        const int cnt = __popc(bal);
        if (j < cnt) return b + __fns(bal, 0, j + 1);
        j -= cnt;
    }
    return -1;
}

// temperature of proposal `it` of a visit that starts at `temp` (see mpp_run_windows): geometric decay per proposal index,
// stopping at the target temperature like the reference's `if T > T_target: T *= alpha`
template <typename R>
__device__ __forceinline__ float visit_temp(const Ctx<R> &c, float temp, int it) {
    if (c.visit_alpha == 1.f) return temp;
    return fmaxf(temp * __powf(c.visit_alpha, (float)it), fminf(temp, c.visit_tfloor));
}

// one record of the per-proposal trace (mpp_window_trace, include/mpp_b200.h); called by one lane
template <typename R>
__device__ __forceinline__ void trace_write(mpp_window_trace *tr, uint32_t flags, int kernel, int n_win, int pid, uint32_t rem_uid, uint32_t add_uid,
                                            const Cand<R> &a, float de, float log_ratio, float temp, uint32_t qk, uint32_t qp, uint32_t qa) {
    mpp_window_trace t;
    t.flags = MPP_TRACE_WRITTEN | flags | ((uint32_t)kernel << 8) | ((uint32_t)min(n_win, 255) << 16) | ((uint32_t)pid << 24);
    t.rem_uid = rem_uid; t.add_uid = add_uid;
    const bool ha = (flags & MPP_TRACE_HAS_ADD) != 0;
    t.add_cls = ha ? a.cls : 0u; t.add_x = a.x; t.add_y = a.y;  // (a rejected move still reports its end point)
    t.add_size = ha ? (float)a.size : 0.f; t.add_ratio = ha ? (float)a.ratio : 0.f; t.add_angle = ha ? (float)a.angle : 0.f;
    t.delta_e = de; t.log_ratio = log_ratio; t.temperature = temp; t.u_accept = u01f(qa);
    t.q[0] = qk; t.q[1] = qp; t.q[2] = qa;
    *tr = t;
}

template <typename R>
struct Eval {  // outcome of evaluating one proposal (warp-uniform)
    int kernel, r;
    int hyp;    // 0: proposed from an empty window (births-only mixture), 1: the reference mixture
    bool has_add, evaluated, accept;
    bool noop;  // accepted proposal that maps the configuration onto itself (see evaluate_proposal): nothing to commit
    Cand<R> a;
};

// Draws and evaluates proposal number `it` of this window's chain against the staged state (read-only).
template <typename R, bool DBG>
__device__ void evaluate_proposal(const Ctx<R> &c, const Ctx<R> &ch, const WinState<R> &w, uint64_t seed, uint32_t win_id, uint64_t sweep_id, int it,
                                  float temp, int lane, R *sx, R *sy, R *po, R *pa, Eval<R> *e, float *dbg_maxdiff, mpp_window_trace *tr) {
    // random words and kernel of proposal `it`: drawn ahead by predraw_births (same counter-based stream)
#ifdef MPP_TRACE
    const long long t_e0 = clock64();
#endif
    const uint4 q0 = make_uint4(w.pq[0][it], w.pq[1][it], w.pq[2][it], w.pq[3][it]);
    const uint4 q1 = make_uint4(w.pq[4][it], w.pq[5][it], w.pq[6][it], w.pq[7][it]);
    const ModelDev &m = c.m;
    const int nc = w.n_win;
    const int hyp = nc > 0 ? 1 : 0;
    e->r = -1; e->has_add = false; e->evaluated = false; e->accept = false; e->noop = false; e->hyp = hyp;
    const int kernel = w.pkern[hyp][it];
    e->kernel = kernel;
    const int wx = w.x1 - w.x0, wy = w.y1 - w.y0;
    int r = -1;
    if (kernel != 0 && kernel != 2) {
        r = w.winlist[min(nc - 1, (int)(u01f(q0.y) * (float)nc))];  // staged indices of the alive window objects, ascending
        if (r < 0) return;  // cannot happen (nc > 0)
    }
    e->r = r;
    Cand<R> &a = e->a;
    if (DBG) { a.x = -1; a.y = -1; }
    float log_ratio = 0.f;  // log(bwd) - log(fwd)
    bool valid = true;
    float pn[3], dm[3];
    int pid_t = 0;  // (trace) mark index of a mark transform
    const uint32_t ruid_t = r >= 0 ? w.uid[r] : 0u, auid_t = w.uid_base + (uint32_t)it;
    const uint32_t rem_t = r >= 0 ? MPP_TRACE_HAS_REM : 0u;
    const float temp_it = visit_temp(c, temp, it);
    switch (kernel) {
    case 0: {  // uniform birth in the window (candidate drawn ahead)
        const float fwd = W2_DIV(pk_of(w, 0, nc), w.lam_unif);
        const float bwd = W2_DIV(pk_of(w, 1, nc + 1), (float)(nc + 1));
        log_ratio = __logf(bwd + W2_EPS) - __logf(fwd + W2_EPS);
        e->has_add = true;
        break;
    }
    case 1: {  // uniform death
        const float fwd = W2_DIV(pk_of(w, 1, nc), (float)nc);
        const float bwd = W2_DIV(pk_of(w, 0, nc - 1), w.lam_unif);
        log_ratio = __logf(bwd + W2_EPS) - __logf(fwd + W2_EPS);
        break;
    }
    case 2: {  // data-driven birth in the window (candidate drawn ahead)
        if (!(w.win_mass > 0.0)) { valid = false; break; }
        const float fwd = W2_DIV(pk_of(w, 2, nc) * dens_of(w, w.pc_detv[hyp][it], w.pc_pn0[hyp][it], w.pc_pn1[hyp][it], w.pc_pn2[hyp][it]), w.lam_data);
        const float bwd = W2_DIV(pk_of(w, 3, nc + 1), (float)(nc + 1));
        log_ratio = __logf(bwd + W2_EPS) - __logf(fwd + W2_EPS);
        e->has_add = true;
        break;
    }
    case 3: {  // data-driven death
        if (!(w.win_mass > 0.0)) { valid = false; break; }
        const float fwd = W2_DIV(pk_of(w, 3, nc), (float)nc);
        const float bwd = W2_DIV(pk_of(w, 2, nc - 1) * dens_of(w, w.detv[r], w.pn0[r], w.pn1[r], w.pn2[r]), w.lam_data);
        log_ratio = __logf(bwd + W2_EPS) - __logf(fwd + W2_EPS);
        break;
    }
    case 4: {  // gaussian translation (symmetric: the proposal densities cancel)
        float d0, d1;
        box_muller_f(q0.z, q0.w, &d0, &d1);
        const int nx_ = min(max((int)((float)w.x[r] + d0 * (float)c.k.trl_sigma), 0), c.H - 1);
        const int ny_ = min(max((int)((float)w.y[r] + d1 * (float)c.k.trl_sigma), 0), c.W - 1);
        a.x = nx_; a.y = ny_;
        if (nx_ < w.x0 || nx_ >= w.x1 || ny_ < w.y0 || ny_ >= w.y1) { valid = false; break; }
        if (nx_ == w.x[r] && ny_ == w.y[r]) { e->noop = true; break; }  // the shift rounds to zero
        a.cls = w.cls[r]; a.size = w.size[r]; a.ratio = w.ratio[r]; a.angle = w.angle[r];
        pixel_info(ch, a.x, a.y, a.cls, lane, &a.detv, pn, dm);
        e->has_add = true;
        break;
    }
    case 5: {  // data-driven translation in the 17x17 window around the object
        const int md = c.k.trl_max_delta;
        const int X0 = max(0, w.x[r] - md), X1 = min(w.x[r] + md + 1, c.H), Y0 = max(0, w.y[r] - md), Y1 = min(w.y[r] + md + 1, c.W);
        const size_t pitch = (size_t)c.W + 1;
        float rs = 0.f;
        if (lane < X1 - X0) rs = (float)(c.rowcum[(size_t)(X0 + lane) * pitch + Y1] - c.rowcum[(size_t)(X0 + lane) * pitch + Y0]);
        float tot_s;
        const int row = warp_pick_ni(rs, u01f(q0.z), lane, &tot_s);
        if (!(tot_s > 0.f)) { valid = false; break; }
        const float dv = lane < Y1 - Y0 ? __ldg(c.det + (size_t)(X0 + row) * c.W + Y0 + lane) : 0.f;
        const int col = warp_pick_ni(dv, u01f(q0.w), lane, nullptr);
        const int ex = X0 + row, ey = Y0 + col;
        a.x = ex; a.y = ey;
        if (ex < w.x0 || ex >= w.x1 || ey < w.y0 || ey >= w.y1) { valid = false; break; }
        if (ex == w.x[r] && ey == w.y[r]) { e->noop = true; break; }  // the object's own pixel was drawn
        a.cls = w.cls[r]; a.size = w.size[r]; a.ratio = w.ratio[r]; a.angle = w.angle[r];
        // backward window (around the end point) + the maps at the end point, one round trip
        const int BX0 = max(0, ex - md), BX1 = min(ex + md + 1, c.H), BY0 = max(0, ey - md), BY1 = min(ey + md + 1, c.W);
        float rb = 0.f;
        if (lane < BX1 - BX0) rb = (float)(c.rowcum[(size_t)(BX0 + lane) * pitch + BY1] - c.rowcum[(size_t)(BX0 + lane) * pitch + BY0]);
        pixel_info(ch, a.x, a.y, a.cls, lane, &a.detv, pn, dm);
        const float tot_e = warp_sum(rb);
        const float fwd = W2_DIV(a.detv, tot_s), bwd = W2_DIV(w.detv[r], tot_e);  // p_kernel / n cancel
        log_ratio = __logf(bwd + W2_EPS) - __logf(fwd + W2_EPS);
        e->has_add = true;
        break;
    }
    default: {  // 6: gaussian mark transform (symmetric), 7: data-driven mark transform
        const int pid = min(2, (int)(u01f(q0.z) * 3.0f));
        pid_t = pid;
        const float v = __ldg(mark_row(c, pid, w.x[r], w.y[r]) + lane);
        float s;
        int ncls;
        R nv;
        const int ocls = cls_of(w.cls[r], pid);
        if (kernel == 6) {
            float d0, d1;
            box_muller_f(q0.w, q1.x, &d0, &d1);
            nv = (pid == 0 ? w.size[r] : (pid == 1 ? w.ratio[r] : w.angle[r])) + (R)(d0 * (float)c.k.trf_sigma[pid]);
            const R vmax = (R)mark_vmax(pid);
            if (pid == 2) { nv = nv - r_floor(nv * (R)(1.0 / 3.14159265358979323846)) * vmax; if (!(nv < vmax) || nv < 0) nv = 0; }
            else nv = r_min(r_max(nv, (R)0), vmax);
            ncls = value_to_class<R>(pid, nv);
            s = warp_sum(v);
        } else {
            ncls = warp_pick_ni(v, u01f(q0.w), lane, &s);
            nv = mark_edge<R>(pid, ncls);
            // the object's own class was drawn and its mark already sits on that class's value: same object
            if (ncls == ocls && nv == (pid == 0 ? w.size[r] : (pid == 1 ? w.ratio[r] : w.angle[r]))) { e->noop = true; break; }
            const float pf = W2_DIV(__shfl_sync(MPP_FULL, v, ncls), s), pb = W2_DIV(__shfl_sync(MPP_FULL, v, ocls), s);
            log_ratio = __logf(pb + W2_EPS) - __logf(pf + W2_EPS);  // p_kernel / n cancel
        }
        const float pnew = __shfl_sync(MPP_FULL, v, ncls);
        a.x = w.x[r]; a.y = w.y[r]; a.size = w.size[r]; a.ratio = w.ratio[r]; a.angle = w.angle[r];
        if (pid == 0) a.size = nv; else if (pid == 1) a.ratio = nv; else a.angle = nv;
        a.cls = (w.cls[r] & ~(0xffu << (8 * pid))) | ((uint32_t)ncls << (8 * pid));
        a.detv = w.detv[r];
        pn[0] = w.pn0[r]; pn[1] = w.pn1[r]; pn[2] = w.pn2[r];
        dm[0] = (float)w.dm0[r]; dm[1] = (float)w.dm1[r]; dm[2] = (float)w.dm2[r];
        pn[pid] = __fdividef(pnew, s);  // (as gather_pixel normalises the staged probabilities)
        dm[pid] = mark_energy_f32(m, pid, pnew);
        e->has_add = true;
        break;
    }
    }
    if (!valid) {
        e->has_add = false;
        if (DBG && tr && lane == 0) trace_write<R>(tr, rem_t | MPP_TRACE_LEFT_WINDOW, kernel, nc, pid_t, ruid_t, auid_t, a, 0.f, 0.f, temp_it, q0.x, q0.y, q1.w);
        return;
    }
    if (e->noop) {
        // The proposal maps the configuration onto itself: Delta-energy 0 and equal forward / backward densities, so the Green
        // ratio is 1 and the reference accepts it (replacing the object by an equal one).  It counts as evaluated and accepted,
        // but there is nothing to commit, and later proposals evaluated against the same state in this round stay valid.
        e->has_add = false; e->evaluated = true; e->accept = true;
        if (DBG && tr && lane == 0)
            trace_write<R>(tr, rem_t | MPP_TRACE_EVALUATED | MPP_TRACE_ACCEPT | MPP_TRACE_IDENTITY, kernel, nc, pid_t, ruid_t, auid_t, a, 0.f, 0.f, temp_it, q0.x, q0.y, q1.w);
        return;
    }
    if (e->has_add) {
        if (r < 0) {  // birth: everything but the Delta-energy was computed ahead
            a.x = w.pc_x[hyp][it]; a.y = w.pc_y[hyp][it]; a.cls = w.pc_cls[hyp][it];
            a.size = w.pc_size[hyp][it]; a.ratio = w.pc_ratio[hyp][it]; a.angle = w.pc_angle[hyp][it];
            a.hl = w.pc_hl[hyp][it]; a.hw = w.pc_hw[hyp][it]; a.ca = w.pc_ca[hyp][it]; a.sa = w.pc_sa[hyp][it];
            a.pos = w.pc_pos[hyp][it]; a.dm0 = w.pc_dm0[hyp][it]; a.dm1 = w.pc_dm1[hyp][it]; a.dm2 = w.pc_dm2[hyp][it];
            a.detv = w.pc_detv[hyp][it]; a.pn0 = w.pc_pn0[hyp][it]; a.pn1 = w.pc_pn1[hyp][it]; a.pn2 = w.pc_pn2[hyp][it];
        } else {
            a.pos = (R)position_energy_f32(a.detv, m.pos_thr);
            a.dm0 = (R)dm[0]; a.dm1 = (R)dm[1]; a.dm2 = (R)dm[2];
            a.pn0 = pn[0]; a.pn1 = pn[1]; a.pn2 = pn[2];
            if (a.size == w.size[r] && a.ratio == w.ratio[r]) {  // a translation or an angle transform keeps the half extents
                a.hl = w.hl[r]; a.hw = w.hw[r];
            } else {
                const R length = r_div_nocheck((R)2 * a.size, (R)1 + a.ratio);
                a.hl = length / (R)2; a.hw = a.ratio * length / (R)2;
            }
            if (a.angle == w.angle[r]) {
                a.ca = w.ca[r]; a.sa = w.sa[r];
            } else {
                float fs, fc;
                __sincosf((float)a.angle, &fs, &fc); a.sa = (R)fs; a.ca = (R)fc;
            }
        }
        // capacity of the destination storage cell (MPP_CELL_CAPACITY slots)
        const int ci = ((a.x >> 5) - w.cx0) * 2 + ((a.y >> 5) - w.cy0);
        uint32_t dmk = w.cmask[ci];
        if (r >= 0 && w.handle[r] != MPP_NO_OBJECT && (int)(w.handle[r] >> 5) == w.ccell[ci]) dmk &= ~(1u << (w.handle[r] & 31));
        if (dmk == 0xffffffffu) {
            e->has_add = false;
            if (DBG && tr && lane == 0) trace_write<R>(tr, rem_t | MPP_TRACE_HAS_ADD | MPP_TRACE_CELL_FULL, kernel, nc, pid_t, ruid_t, auid_t, a, 0.f, 0.f, temp_it, q0.x, q0.y, q1.w);
            return;
        }
    }
#ifdef MPP_TRACE
    const long long t_d0 = clock64();
#endif
    const R de = delta_staged(m, w, r, e->has_add, a, lane, sx, sy, po, pa);
#ifdef MPP_TRACE
    if (lane == 0) {
        const long long t_d1 = clock64();
        atomicAdd(c.kstats + 40 + kernel, (unsigned long long)(t_d0 - t_e0));   // draw part of this kernel
        atomicAdd(c.kstats + 48 + kernel, (unsigned long long)(t_d1 - t_d0));   // Delta-energy part
        atomicAdd(c.kstats + 56 + kernel, 1ull);
    }
#endif
#ifndef MPP_TRACE
    if (DBG && dbg_maxdiff) {
        const R db = delta_brute(m, w, r, e->has_add, a, lane, sx, sy);
        const float diff = fabsf((float)(de - db));
        if (lane == 0) atomicMax(reinterpret_cast<int *>(dbg_maxdiff), __float_as_int(diff));
    }
#endif
    const float la = W2_DIV(-(float)de, temp_it) + log_ratio;
    e->evaluated = true;
    e->accept = __logf(u01f(q1.w) + W2_EPS) < la;
    if (DBG && tr && lane == 0)
        trace_write<R>(tr, rem_t | (e->has_add ? MPP_TRACE_HAS_ADD : 0u) | MPP_TRACE_EVALUATED | (e->accept ? MPP_TRACE_ACCEPT : 0u), kernel, nc, pid_t, ruid_t,
                       auid_t, a, (float)de, log_ratio, temp_it, q0.x, q0.y, q1.w);
}

// Applies an accepted proposal to the staged state and writes the new record (executed by the warp that evaluated it:
// its po / pa scratch still holds the pair values between every staged entry and the added object).  The occupancy
// masks of the window's storage cells are only published at the end of the visit.
template <typename R> __device__ __forceinline__ void rebuild_winlist(WinState<R> &w, int lane);

template <typename R, bool SPLIT = false>
__device__ void commit_proposal(const Ctx<R> &c, const Ctx<R> &ch, WinState<R> &w, const Eval<R> &e, int it, int lane, R *sx, R *sy, const R *po, const R *pa) {
    const ModelDev &m = c.m;
    const int r = e.r;
#ifdef MPP_TRACE
    const long long tc0 = clock64();  // (instrumented build: phases of a commit, tools/visit_timers.py)
#endif
    if (r >= 0 && lane == 0) {
        w.flags[r] = 0;
        const uint32_t h = w.handle[r];
        for (int q = 0; q < 4; ++q)
            if (w.ccell[q] == (int)(h >> 5)) w.cmask[q] &= ~(1u << (h & 31));
        w.n_win -= 1; w.dn -= 1; w.masks_dirty = 1;
    }
    int s = -1;
    if (e.has_add) {
        // free staged slot: the removed entry if any, else the first dead entry, else append
        s = r;
        if (s < 0) {
            const int n = w.n;
#pragma unroll 1
            for (int b = 0; b < n && s < 0; b += 32) {
                const uint32_t bal = __ballot_sync(MPP_FULL, b + lane < n && !(w.flags[b + lane] & W2_ALIVE));
                if (bal) s = b + __ffs(bal) - 1;
            }
            if (s < 0) s = n;  // n < W2_K is guaranteed by the caller
        }
        // everything the new entry needs is computed from the candidate's registers (warp-uniform); lane 0 then only stores
        const Cand<R> &a = e.a;
        R tm0, tm1, tm2, fa, fg;
        shape_terms<R>(m, a.dm0, a.dm1, a.dm2, &tm0, &tm1, &tm2);
        unit_form_vals<R>(m, tm0, tm1, tm2, a.hl, a.hw, a.ratio, a.pos, &fa, &fg);
        const R rad = r_sqrt_fast(a.hl * a.hl + a.hw * a.hw);
        const uint32_t uid = w.uid_base + (uint32_t)it;
        if (lane == 0) {
            if (s == w.n) w.n = s + 1;
            const int ci = ((a.x >> 5) - w.cx0) * 2 + ((a.y >> 5) - w.cy0);
            const uint32_t cm = w.cmask[ci];
            const int slot = __ffs(~cm) - 1;
            const uint32_t h = (uint32_t)w.ccell[ci] * 32u + slot;
            Rec<R> rec;
            rec.x = a.x; rec.y = a.y; rec.cls = a.cls; rec.uid = uid;
            rec.size = a.size; rec.ratio = a.ratio; rec.angle = a.angle;
            rec.e_pos = a.pos; rec.e_m[0] = tm0; rec.e_m[1] = tm1; rec.e_m[2] = tm2;
            rec.hl = a.hl; rec.hw = a.hw; rec.ca = a.ca; rec.sa = a.sa; rec.pad = 0;
            store_rec(rec_ptr<SPLIT>(c, h), rec);
            w.cmask[ci] = cm | (1u << slot);
            w.x[s] = a.x; w.y[s] = a.y; w.cls[s] = a.cls; w.handle[s] = h; w.uid[s] = uid;
            w.size[s] = a.size; w.ratio[s] = a.ratio; w.angle[s] = a.angle;
            w.hl[s] = a.hl; w.hw[s] = a.hw; w.ca[s] = a.ca; w.sa[s] = a.sa; w.rad[s] = rad;
            w.pos[s] = a.pos; w.dm0[s] = a.dm0; w.dm1[s] = a.dm1; w.dm2[s] = a.dm2;
            w.tm0[s] = tm0; w.tm1[s] = tm1; w.tm2[s] = tm2; w.fa[s] = fa; w.fg[s] = fg;
            w.detv[s] = a.detv; w.pn0[s] = a.pn0; w.pn1[s] = a.pn1; w.pn2[s] = a.pn2;
            w.flags[s] = W2_ALIVE | W2_WIN | W2_INNER;
            w.n_win += 1; w.dn += 1; w.masks_dirty = 1;
        }
    }
    __syncwarp();
    if ((r >= 0) != (s >= 0)) rebuild_winlist(w, lane);  // a birth or a death changes the list the proposals pick from
    // reductions: incremental for the addition (pair values are in po / pa); entries whose best partner was the
    // removed object are rescanned; the new object's own top-2 is the top-2 of po / pa
    const int n = w.n;
    R n_o1 = 0, n_o2 = 0, n_a1 = 0, n_a2 = 0;
    int n_ao = -1, n_aa = -1, n_ao2 = -1, n_aa2 = -1;
    bool any_pair = false;
#ifdef MPP_TRACE
    const long long tc1 = clock64();
#endif
#pragma unroll 1
    for (int k = lane; k < n; k += 32) {
        if (k == s || !(w.flags[k] & W2_ALIVE)) continue;
        // the removed object was this entry's best or second-best partner: its top-2 must be rescanned -- unless the object is
        // REPLACED in the same staged slot (a move: s == r) and its new pair value keeps its place: then only the value changes.
        // (Most accepted moves are mark transforms that leave the alignment value with every neighbour as it was.)
        bool redo_ov = r >= 0 && (w.aov[k] == r || w.aov2[k] == r), redo_al = r >= 0 && (w.aal[k] == r || w.aal2[k] == r);
        const bool was_ov = redo_ov && s == r, was_al = redo_al && s == r;  // r held a place in this entry's top-2 and is being replaced
        const R o1_old = w.ov1[k], a1_old = w.al1[k];
        const R o = s >= 0 ? po[k] : (R)0, al = s >= 0 ? pa[k] : (R)0;
        if (s >= 0 && s == r && (w.flags[k] & W2_INNER)) {
            if (redo_ov) {
                if (w.aov[k] == r) {  // was the best partner
                    if (o > (R)0 && o >= w.ov2[k]) { w.ov1[k] = o; redo_ov = false; }
                    else if (!(o > (R)0) && !(w.ov2[k] > (R)0)) { w.ov1[k] = 0; w.aov[k] = -1; redo_ov = false; }
                } else {  // was the second best
                    if (o > w.ov1[k]) { const R t1 = w.ov1[k]; const short a1 = w.aov[k]; w.ov1[k] = o; w.aov[k] = (short)s; w.ov2[k] = t1; w.aov2[k] = a1; redo_ov = false; }
                    else if (o > (R)0 && o >= w.ov2[k]) { w.ov2[k] = o; redo_ov = false; }
                }
            }
            if (redo_al) {
                if (w.aal[k] == r) {
                    if (al > (R)0 && al >= w.al2[k]) { w.al1[k] = al; redo_al = false; }
                    else if (!(al > (R)0) && !(w.al2[k] > (R)0)) { w.al1[k] = 0; w.aal[k] = -1; redo_al = false; }
                } else {
                    if (al > w.al1[k]) { const R t1 = w.al1[k]; const short a1 = w.aal[k]; w.al1[k] = al; w.aal[k] = (short)s; w.al2[k] = t1; w.aal2[k] = a1; redo_al = false; }
                    else if (al > (R)0 && al >= w.al2[k]) { w.al2[k] = al; redo_al = false; }
                }
            }
        }
        if (s >= 0) {
            if (o > (R)0 || al > (R)0) {  // isolated objects (the common case) skip all of this
                any_pair = true;
                if (o > n_o1) { n_o2 = n_o1; n_ao2 = n_ao; n_o1 = o; n_ao = k; } else if (o > n_o2) { n_o2 = o; n_ao2 = k; }
                if (al > n_a1) { n_a2 = n_a1; n_aa2 = n_aa; n_a1 = al; n_aa = k; } else if (al > n_a2) { n_a2 = al; n_aa2 = k; }
                if (!redo_ov && !was_ov && (w.flags[k] & W2_INNER)) {
                    if (o > w.ov1[k]) { w.ov2[k] = w.ov1[k]; w.aov2[k] = w.aov[k]; w.ov1[k] = o; w.aov[k] = (short)s; }
                    else if (o > w.ov2[k]) { w.ov2[k] = o; w.aov2[k] = (short)s; }
                }
                if (!redo_al && !was_al && (w.flags[k] & W2_INNER)) {
                    if (al > w.al1[k]) { w.al2[k] = w.al1[k]; w.aal2[k] = w.aal[k]; w.al1[k] = al; w.aal[k] = (short)s; }
                    else if (al > w.al2[k]) { w.al2[k] = al; w.aal2[k] = (short)s; }
                }
            }
        }
        if ((redo_ov || redo_al) && (w.flags[k] & W2_INNER)) recompute_top2(ch.m, w, k, redo_ov, redo_al, sx, sy);
        if ((w.flags[k] & W2_INNER) && (w.ov1[k] != o1_old || w.al1[k] != a1_old)) w.fcur[k] = f_obj(m, w, k, w.ov1[k], w.al1[k]);
    }
#ifdef MPP_TRACE
    __syncwarp();
    const long long tc2 = clock64();
    if (lane == 0) { atomicAdd(c.kstats + 35, (unsigned long long)(tc1 - tc0)); atomicAdd(c.kstats + 36, (unsigned long long)(tc2 - tc1)); atomicAdd(c.kstats + 38, 1ull); }
#endif
    if (s >= 0 && !__any_sync(MPP_FULL, any_pair)) {  // no partner within reach of the new object
        if (lane == 0) {
            w.ov1[s] = 0; w.ov2[s] = 0; w.al1[s] = 0; w.al2[s] = 0; w.aov[s] = -1; w.aov2[s] = -1; w.aal[s] = -1; w.aal2[s] = -1;
            w.fcur[s] = f_obj(m, w, s, (R)0, (R)0);
        }
        __syncwarp();
#ifdef MPP_TRACE
        if (lane == 0) atomicAdd(c.kstats + 37, (unsigned long long)(clock64() - tc2));
#endif
        return;
    }
    if (s >= 0) {  // merge the per-lane top-2 of (po, pa) into the new object's reductions (ties: lowest lane first)
        const R mo = warp_max_nonneg(n_o1);
        const int lo = __ffs(__ballot_sync(MPP_FULL, n_o1 == mo)) - 1;
        const R co = lane == lo ? n_o2 : n_o1;
        const R so = warp_max_nonneg(co);
        const int lo2 = __ffs(__ballot_sync(MPP_FULL, co == so)) - 1;
        const int ao = __shfl_sync(MPP_FULL, n_ao, lo);
        const int ao2 = __shfl_sync(MPP_FULL, lane == lo ? n_ao2 : n_ao, lo2);
        const R ma = warp_max_nonneg(n_a1);
        const int la = __ffs(__ballot_sync(MPP_FULL, n_a1 == ma)) - 1;
        const R ca2 = lane == la ? n_a2 : n_a1;
        const R sa2 = warp_max_nonneg(ca2);
        const int la2 = __ffs(__ballot_sync(MPP_FULL, ca2 == sa2)) - 1;
        const int aa = __shfl_sync(MPP_FULL, n_aa, la);
        const int aa2 = __shfl_sync(MPP_FULL, lane == la ? n_aa2 : n_aa, la2);
        if (lane == 0) {
            w.ov1[s] = mo; w.ov2[s] = so; w.aov[s] = (short)(mo > (R)0 ? ao : -1); w.aov2[s] = (short)(so > (R)0 ? ao2 : -1);
            w.al1[s] = ma; w.al2[s] = sa2; w.aal[s] = (short)(ma > (R)0 ? aa : -1); w.aal2[s] = (short)(sa2 > (R)0 ? aa2 : -1);
            w.fcur[s] = f_obj(m, w, s, mo, ma);
        }
    }
    __syncwarp();
#ifdef MPP_TRACE
    if (lane == 0) atomicAdd(c.kstats + 37, (unsigned long long)(clock64() - tc2));
#endif
}

// ================================================================================================ SIMT mode
// Lane-per-proposal evaluation: ONE warp per window evaluates 32 consecutive proposals of the chain at once, each lane
// running a whole proposal (draw, map gathers, Delta-energy over the staged objects, Green ratio) by itself.  The first
// accepted lane is committed (warp-cooperatively), the later ones are discarded and re-drawn: same "speculative moves"
// semantics as the warp-per-proposal mode, but a window now costs one warp instead of up to eight, every instruction is
// fetched once for 32 proposals, and the loops over the staged objects read shared memory by broadcast.

// inverse-CDF pick over n <= 32 consecutive floats in global memory: first k with cumulative sum > target.  All the loads
// are issued before the scan (independent loads, one memory round trip) instead of one dependent load per step.
__device__ __forceinline__ int scan32(const float *v, int n, float target, float *picked) {
    float acc = 0.f, lastv = 0.f;
    int last = 0;
    bool done = false;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const float x = v[k];
        if (!done && k < n && x > 0.f) { last = k; lastv = x; acc += x; done = acc > target; }
    }
    *picked = lastv;
    return last;
}
__device__ MPP_SCAN_INL int scan_pick_global(const float *row, int n, float target, float *picked) {
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < n ? __ldg(row + k) : 0.f;
    return scan32(v, n, target, picked);
}
// ... over one 128-byte aligned mark row (32 classes): eight 16-byte loads
__device__ MPP_SCAN_INL int scan_pick_row(const float *row, float target, float *picked) {
    float v[32];
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float4 q = __ldg(r4 + k); v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w; }
    return scan32(v, 32, target, picked);
}

// Proposals drawn ahead (warp mode).  A birth (uniform or data-driven) does not depend on the configuration: its position,
// marks, map values and unit energies are functions of the random words and of the maps only.  One warp therefore draws
// the births of a whole visit at once, one proposal per lane (every memory round trip is shared by up to 32 proposals),
// while warp 0 is still staging the neighbourhood; the speculative rounds then only compute Delta-energies.  `hyp`
// selects the kernel mixture (0: empty window, births only; 1: the reference mixture): both are drawn, the rounds
// pick the one that matches the window's state at that point of the chain.  `rowm` is 32 floats of warp-private
// shared memory.  Must be called by a full warp.
template <typename R>
__device__ __noinline__ void predraw_births(const Ctx<R> &c, WinState<R> &w, int hyp, int chunk, int per_visit, int x0, int x1, int y0, int y1,
                                            uint64_t seed, uint32_t win_id, uint64_t sweep_id, R *rowm, int lane) {
    const ModelDev &m = c.m;
    const int wx = x1 - x0, wy = y1 - y0;
    // detection mass of the window rows (same arithmetic as the staging of row_mass / win_mass)
    double rm = 0.0;
    {
        const size_t pitch = (size_t)c.W + 1;
        if (lane < wx) rm = c.rowcum[(size_t)(x0 + lane) * pitch + y1] - c.rowcum[(size_t)(x0 + lane) * pitch + y0];
    }
    const double win_mass = warp_sum(rm);
    rowm[lane] = (R)(float)rm;
    __syncwarp();
    const int it = chunk * 32 + lane;
    if (it < per_visit) {
        Philox rng(seed, win_id, (uint32_t)sweep_id, ((uint32_t)(sweep_id >> 32) << 20) ^ (uint32_t)it ^ 0x77000000u);
        const uint4 q0 = rng.next(), q1 = rng.next();
        if (hyp == 0) {
            w.pq[0][it] = q0.x; w.pq[1][it] = q0.y; w.pq[2][it] = q0.z; w.pq[3][it] = q0.w;
            w.pq[4][it] = q1.x; w.pq[5][it] = q1.y; w.pq[6][it] = q1.z; w.pq[7][it] = q1.w;
        }
        int kernel;
        {
            const float uk = u01f(q0.x);
            if (hyp) { float acc = 0; kernel = 7; for (int k = 0; k < 7; ++k) { acc += c.k.pf[k]; if (uk < acc) { kernel = k; break; } } }
            else kernel = uk < c.k.pk_e0 ? 0 : 2;
        }
        w.pkern[hyp][it] = (unsigned char)kernel;
        int x = 0, y = 0;
        uint32_t cls = 0;
        R size = 0, ratio = 0, angle = 0;
        bool birth = false;
        if (kernel == 0) {  // uniform birth in the window
            x = x0 + min(wx - 1, (int)(u01f(q0.z) * (float)wx));
            y = y0 + min(wy - 1, (int)(u01f(q0.w) * (float)wy));
            size = (R)(u01f(q1.x) * 32.0f); ratio = (R)u01f(q1.y); angle = (R)(u01f(q1.z) * 3.14159265358979f);
            cls = pack_cls(value_to_class<R>(0, size), value_to_class<R>(1, ratio), value_to_class<R>(2, angle));
            birth = true;
        } else if (kernel == 2 && win_mass > 0.0) {  // data-driven birth: pixel ~ detection map, marks ~ mark maps at the pixel
            float acc = 0.f;
            const float tr = u01f(q0.z) * (float)win_mass;
            int row = 0;
            for (int k = 0; k < wx; ++k) { const float v = (float)rowm[k]; if (v > 0.f) { row = k; acc += v; if (acc > tr) break; } }
            float dv;
            const int col = scan_pick_global(c.det + (size_t)(x0 + row) * c.W + y0, wy, u01f(q0.w) * (float)rowm[row], &dv);
            x = x0 + row; y = y0 + col;
            const size_t pix = (size_t)x * c.W + y, plane = (size_t)c.H * c.W;
            float pv;
            const int c0 = scan_pick_row(mark_row(c, 0, x, y), u01f(q1.x) * __ldg(c.marksum + pix), &pv);
            const int c1 = scan_pick_row(mark_row(c, 1, x, y), u01f(q1.y) * __ldg(c.marksum + plane + pix), &pv);
            const int c2 = scan_pick_row(mark_row(c, 2, x, y), u01f(q1.z) * __ldg(c.marksum + 2 * plane + pix), &pv);
            cls = pack_cls(c0, c1, c2);
            size = mark_edge<R>(0, c0); ratio = mark_edge<R>(1, c1); angle = mark_edge<R>(2, c2);
            birth = true;
        }
        if (birth) {
            float detv, pn[3], dm[3];
            gather_pixel(c, x, y, cls, &detv, pn, dm);
            const R length = r_div_nocheck((R)2 * size, (R)1 + ratio);
            float fs, fc;
            __sincosf((float)angle, &fs, &fc);
            w.pc_x[hyp][it] = x; w.pc_y[hyp][it] = y; w.pc_cls[hyp][it] = cls;
            w.pc_size[hyp][it] = size; w.pc_ratio[hyp][it] = ratio; w.pc_angle[hyp][it] = angle;
            w.pc_hl[hyp][it] = length / (R)2; w.pc_hw[hyp][it] = ratio * length / (R)2; w.pc_ca[hyp][it] = (R)fc; w.pc_sa[hyp][it] = (R)fs;
            w.pc_pos[hyp][it] = (R)position_energy_f32(detv, m.pos_thr);
            w.pc_dm0[hyp][it] = (R)dm[0]; w.pc_dm1[hyp][it] = (R)dm[1]; w.pc_dm2[hyp][it] = (R)dm[2];
            w.pc_detv[hyp][it] = detv; w.pc_pn0[hyp][it] = pn[0]; w.pc_pn1[hyp][it] = pn[1]; w.pc_pn2[hyp][it] = pn[2];
        }
    }
    __syncwarp();
}

// Delta-energy of removing staged entry r and/or adding `a`, computed by ONE thread over the staged state
template <typename R>
__device__ __forceinline__ R delta_lane(const ModelDev &m, const WinState<R> &w, int r, bool has_add, const Cand<R> &a, R *sx, R *sy) {
    Geo<R> ga; ga.x = a.x; ga.y = a.y; ga.hl = a.hl; ga.hw = a.hw; ga.ca = a.ca; ga.sa = a.sa;
    const R rad_a = has_add ? r_sqrt_fast(a.hl * a.hl + a.hw * a.hw) : (R)0;
    const int rx = r >= 0 ? w.x[r] : 0, ry = r >= 0 ? w.y[r] : 0;
    R acc = 0, ov_add = 0, al_add = 0;
    const int n = w.n;
    for (int k = 0; k < n; ++k) {
        if (k == r || !(w.flags[k] & W2_ALIVE)) continue;
        bool touched = false;
        R ov_a = w.ov1[k], al_a = w.al1[k];
        const int kx = w.x[k], ky = w.y[k];
        if (r >= 0) {
            const int dx = kx - rx, dy = ky - ry;
            if (dx * dx + dy * dy <= m.max_d2) {
                touched = true;
                if (w.aov[k] == r) ov_a = w.ov2[k];
                if (w.aal[k] == r) al_a = w.al2[k];
            }
        }
        if (has_add) {
            const int dx = kx - a.x, dy = ky - a.y, d2 = dx * dx + dy * dy;
            if (d2 <= m.max_d2) {
                touched = true;
                if (d2 <= m.ov_d2) { const R o = pair_ov_w(m, w, k, ga, rad_a, d2, sx, sy); ov_a = r_max(ov_a, o); ov_add = r_max(ov_add, o); }
                if (d2 <= m.al_d2) { const R al = align_magnitude(geo_w(w, k), ga, m.rewarding); al_a = r_max(al_a, al); al_add = r_max(al_add, al); }
            }
        }
        if (touched) acc += f_obj(m, w, k, ov_a, al_a) - f_obj(m, w, k, w.ov1[k], w.al1[k]);
    }
    if (has_add) {
        Terms<R> t;
        t.pos = a.pos;
        shape_terms<R>(m, a.dm0, a.dm1, a.dm2, &t.m0, &t.m1, &t.m2);
        t.ov = ov_add; t.al = (m.rewarding ? (R)-1 : (R)1) * al_add;
        t.area = area_prior_fast<R>(m, a.hl, a.hw);
        t.ratio = r_abs((R)m.f_target_ratio - a.ratio);
        acc += combine_fast(m, t);
    }
    if (r >= 0) acc -= f_obj(m, w, r, w.ov1[r], w.al1[r]);
    return acc;
}

// brute-force twin of delta_lane (debug): every reduction recomputed from scratch by the same thread
template <typename R>
__device__ R delta_lane_brute(const ModelDev &m, const WinState<R> &w, int r, bool has_add, const Cand<R> &a, R *sx, R *sy) {
    Geo<R> ga; ga.x = a.x; ga.y = a.y; ga.hl = a.hl; ga.hw = a.hw; ga.ca = a.ca; ga.sa = a.sa;
    R acc = 0, oadd = 0, aadd = 0;
    const int n = w.n;
    for (int k = 0; k < n; ++k) {
        if (!(w.flags[k] & W2_ALIVE)) continue;
        R ob = 0, ab = 0, oa = 0, aa = 0;
        const Geo<R> gk = geo_w(w, k);
        for (int v = 0; v < n; ++v) {
            if (v == k || !(w.flags[v] & W2_ALIVE)) continue;
            const int dx = w.x[v] - gk.x, dy = w.y[v] - gk.y, d2 = dx * dx + dy * dy;
            if (d2 > m.max_d2) continue;
            R o = 0, al = 0;
            if (d2 <= m.ov_d2) o = pair_overlap(m, gk, geo_w(w, v), d2, sx, sy);
            if (d2 <= m.al_d2) al = align_magnitude(gk, geo_w(w, v), m.rewarding);
            ob = r_max(ob, o); ab = r_max(ab, al);
            if (v != r) { oa = r_max(oa, o); aa = r_max(aa, al); }
        }
        if (has_add && k != r) {
            const int dx = a.x - gk.x, dy = a.y - gk.y, d2 = dx * dx + dy * dy;
            if (d2 <= m.ov_d2) { const R o = pair_overlap(m, gk, ga, d2, sx, sy); oa = r_max(oa, o); oadd = r_max(oadd, o); }
            if (d2 <= m.al_d2) { const R al = align_magnitude(gk, ga, m.rewarding); aa = r_max(aa, al); aadd = r_max(aadd, al); }
        }
        if (k == r) acc -= f_obj(m, w, k, ob, ab);
        else if (oa != ob || aa != ab) acc += f_obj(m, w, k, oa, aa) - f_obj(m, w, k, ob, ab);
    }
    if (has_add) {
        Terms<R> t;
        t.pos = a.pos;
        shape_terms<R>(m, a.dm0, a.dm1, a.dm2, &t.m0, &t.m1, &t.m2);
        t.ov = oadd; t.al = (m.rewarding ? (R)-1 : (R)1) * aadd;
        t.area = area_prior_fast<R>(m, a.hl, a.hw);
        t.ratio = r_abs((R)m.f_target_ratio - a.ratio);
        acc += combine_fast(m, t);
    }
    return acc;
}

// Proposal number `it` of this window's chain, drawn and evaluated by the calling THREAD against the staged state.
template <typename R, bool DBG>
__device__ __forceinline__ void evaluate_lane(const Ctx<R> &c, const WinState<R> &w, uint64_t seed, uint32_t win_id, uint64_t sweep_id, int it, float temp,
                              R *sx, R *sy, Eval<R> *e, float *dbg_maxdiff) {
    Philox rng(seed, win_id, (uint32_t)sweep_id, ((uint32_t)(sweep_id >> 32) << 20) ^ (uint32_t)it ^ 0x51000000u);
    const uint4 q0 = rng.next(), q1 = rng.next();
    const ModelDev &m = c.m;
    const int nc = w.n_win;
    e->r = -1; e->has_add = false; e->evaluated = false; e->accept = false;
    int kernel;
    {
        const float uk = u01f(q0.x);
        if (nc > 0) { float acc = 0; kernel = 7; for (int k = 0; k < 7; ++k) { acc += w.pkf[k]; if (uk < acc) { kernel = k; break; } } }
        else kernel = uk < w.pk_e0 ? 0 : 2;
    }
    e->kernel = kernel;
    const int wx = w.x1 - w.x0, wy = w.y1 - w.y0;
    int r = -1;
    if (kernel != 0 && kernel != 2) r = w.winlist[min(nc - 1, (int)(u01f(q0.y) * (float)nc))];
    e->r = r;
    Cand<R> &a = e->a;
    bool valid = true, has_add = false, need_gather = false;
    float log_ratio = 0.f, aux0 = 0.f, aux1 = 0.f;  // kernel-specific proposal densities gathered in phase 1
    int pid = 0;
    // ---- phase 1 (divergent): where / what is proposed
    switch (kernel) {
    case 0:
        a.x = w.x0 + min(wx - 1, (int)(u01f(q0.z) * (float)wx));
        a.y = w.y0 + min(wy - 1, (int)(u01f(q0.w) * (float)wy));
        a.size = (R)(u01f(q1.x) * 32.0f); a.ratio = (R)u01f(q1.y); a.angle = (R)(u01f(q1.z) * 3.14159265358979f);
        a.cls = pack_cls(value_to_class<R>(0, a.size), value_to_class<R>(1, a.ratio), value_to_class<R>(2, a.angle));
        has_add = true; need_gather = true;
        break;
    case 1:
        break;
    case 2: {
        if (!(w.win_mass > 0.0)) { valid = false; break; }
        float acc = 0.f;
        const float tr = u01f(q0.z) * (float)w.win_mass;
        int row = 0;
        for (int k = 0; k < wx; ++k) { const float v = w.row_mass[k]; if (v > 0.f) { row = k; acc += v; if (acc > tr) break; } }
        float dv;
        const int col = scan_pick_global(c.det + (size_t)(w.x0 + row) * c.W + w.y0, wy, u01f(q0.w) * w.row_mass[row], &dv);
        a.x = w.x0 + row; a.y = w.y0 + col;
        const size_t pix = (size_t)a.x * c.W + a.y, plane = (size_t)c.H * c.W;
        float pv;
        const int c0 = scan_pick_row(mark_row(c, 0, a.x, a.y), u01f(q1.x) * __ldg(c.marksum + pix), &pv);
        const int c1 = scan_pick_row(mark_row(c, 1, a.x, a.y), u01f(q1.y) * __ldg(c.marksum + plane + pix), &pv);
        const int c2 = scan_pick_row(mark_row(c, 2, a.x, a.y), u01f(q1.z) * __ldg(c.marksum + 2 * plane + pix), &pv);
        a.cls = pack_cls(c0, c1, c2);
        a.size = mark_edge<R>(0, c0); a.ratio = mark_edge<R>(1, c1); a.angle = mark_edge<R>(2, c2);
        has_add = true; need_gather = true;
        break;
    }
    case 3:
        if (!(w.win_mass > 0.0)) valid = false;
        break;
    case 4: {
        float d0, d1;
        box_muller_f(q0.z, q0.w, &d0, &d1);
        const int nx_ = min(max((int)((float)w.x[r] + d0 * (float)c.k.trl_sigma), 0), c.H - 1);
        const int ny_ = min(max((int)((float)w.y[r] + d1 * (float)c.k.trl_sigma), 0), c.W - 1);
        if (nx_ < w.x0 || nx_ >= w.x1 || ny_ < w.y0 || ny_ >= w.y1) { valid = false; break; }
        a.x = nx_; a.y = ny_; a.cls = w.cls[r]; a.size = w.size[r]; a.ratio = w.ratio[r]; a.angle = w.angle[r];
        has_add = true; need_gather = true;
        break;
    }
    case 5: {
        const int md = c.k.trl_max_delta;
        const int X0 = max(0, w.x[r] - md), X1 = min(w.x[r] + md + 1, c.H), Y0 = max(0, w.y[r] - md), Y1 = min(w.y[r] + md + 1, c.W);
        const size_t pitch = (size_t)c.W + 1;
        float rs[17], tot_s = 0.f;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            rs[i] = (i < X1 - X0) ? (float)(c.rowcum[(size_t)(X0 + i) * pitch + Y1] - c.rowcum[(size_t)(X0 + i) * pitch + Y0]) : 0.f;
            tot_s += rs[i];
        }
        if (!(tot_s > 0.f)) { valid = false; break; }
        int row = 0;
        {
            float acc = 0.f;
            const float tr = u01f(q0.z) * tot_s;
#pragma unroll
            for (int i = 0; i < 17; ++i) if (rs[i] > 0.f && acc <= tr) { row = i; acc += rs[i]; }
        }
        float dv;
        const float rowtot = (float)(c.rowcum[(size_t)(X0 + row) * pitch + Y1] - c.rowcum[(size_t)(X0 + row) * pitch + Y0]);
        const int col = scan_pick_global(c.det + (size_t)(X0 + row) * c.W + Y0, Y1 - Y0, u01f(q0.w) * rowtot, &dv);
        const int ex = X0 + row, ey = Y0 + col;
        if (ex < w.x0 || ex >= w.x1 || ey < w.y0 || ey >= w.y1) { valid = false; break; }
        a.x = ex; a.y = ey; a.cls = w.cls[r]; a.size = w.size[r]; a.ratio = w.ratio[r]; a.angle = w.angle[r];
        const int BX0 = max(0, ex - md), BX1 = min(ex + md + 1, c.H), BY0 = max(0, ey - md), BY1 = min(ey + md + 1, c.W);
        float tot_e = 0.f;
#pragma unroll
        for (int i = 0; i < 17; ++i)
            if (i < BX1 - BX0) tot_e += (float)(c.rowcum[(size_t)(BX0 + i) * pitch + BY1] - c.rowcum[(size_t)(BX0 + i) * pitch + BY0]);
        aux0 = tot_s; aux1 = tot_e;
        has_add = true; need_gather = true;
        break;
    }
    default: {  // 6: gaussian mark transform, 7: data-driven mark transform
        pid = min(2, (int)(u01f(q0.z) * 3.0f));
        const size_t pix = (size_t)w.x[r] * c.W + w.y[r], plane = (size_t)c.H * c.W;
        const float *row = mark_row(c, pid, w.x[r], w.y[r]);
        const float s = __ldg(c.marksum + (size_t)pid * plane + pix);
        const int ocls = cls_of(w.cls[r], pid);
        int ncls;
        R nv;
        float pnew;
        if (kernel == 6) {
            float d0, d1;
            box_muller_f(q0.w, q1.x, &d0, &d1);
            nv = (pid == 0 ? w.size[r] : (pid == 1 ? w.ratio[r] : w.angle[r])) + (R)(d0 * (float)c.k.trf_sigma[pid]);
            const R vmax = (R)mark_vmax(pid);
            if (pid == 2) { nv = nv - r_floor(nv / vmax) * vmax; if (!(nv < vmax) || nv < 0) nv = 0; }
            else nv = r_min(r_max(nv, (R)0), vmax);
            ncls = value_to_class<R>(pid, nv);
            pnew = __ldg(row + ncls);
        } else {
            ncls = scan_pick_row(row, u01f(q0.w) * s, &pnew);
            nv = mark_edge<R>(pid, ncls);
            const float pb = __ldg(row + ocls);
            log_ratio = __logf(__fdividef(pb, s) + W2_EPS) - __logf(__fdividef(pnew, s) + W2_EPS);  // p_kernel / n cancel
        }
        a.x = w.x[r]; a.y = w.y[r]; a.size = w.size[r]; a.ratio = w.ratio[r]; a.angle = w.angle[r];
        if (pid == 0) a.size = nv; else if (pid == 1) a.ratio = nv; else a.angle = nv;
        a.cls = (w.cls[r] & ~(0xffu << (8 * pid))) | ((uint32_t)ncls << (8 * pid));
        a.detv = w.detv[r];
        a.pn0 = w.pn0[r]; a.pn1 = w.pn1[r]; a.pn2 = w.pn2[r];
        a.dm0 = w.dm0[r]; a.dm1 = w.dm1[r]; a.dm2 = w.dm2[r];
        const float pnn = __fdividef(pnew, s);
        const R dmn = (R)mark_energy_f32(m, pid, pnew);
        if (pid == 0) { a.pn0 = pnn; a.dm0 = dmn; } else if (pid == 1) { a.pn1 = pnn; a.dm1 = dmn; } else { a.pn2 = pnn; a.dm2 = dmn; }
        has_add = true;
        break;
    }
    }
    if (!valid) return;
    // ---- phase 2 (common): maps at the proposed pixel, geometry
    if (need_gather) {
        float pn[3], dm[3];
        gather_pixel(c, a.x, a.y, a.cls, &a.detv, pn, dm);
        a.pn0 = pn[0]; a.pn1 = pn[1]; a.pn2 = pn[2];
        a.dm0 = (R)dm[0]; a.dm1 = (R)dm[1]; a.dm2 = (R)dm[2];
    }
    if (has_add) {
        a.pos = (R)position_energy_f32(a.detv, m.pos_thr);
        const R length = ((R)2 * a.size) / ((R)1 + a.ratio);
        a.hl = length / (R)2; a.hw = a.ratio * length / (R)2;
        if (r >= 0 && a.angle == w.angle[r]) { a.ca = w.ca[r]; a.sa = w.sa[r]; }
        else { float fs, fc; __sincosf((float)a.angle, &fs, &fc); a.sa = (R)fs; a.ca = (R)fc; }
        const int ci = ((a.x >> 5) - w.cx0) * 2 + ((a.y >> 5) - w.cy0);
        uint32_t dmk = w.cmask[ci];
        if (r >= 0 && (int)(w.handle[r] >> 5) == w.ccell[ci]) dmk &= ~(1u << (w.handle[r] & 31));
        if (dmk == 0xffffffffu) return;
    }
    // ---- phase 3: proposal ratio log(bwd) - log(fwd)
    switch (kernel) {
    case 0: log_ratio = __logf(pk_of(w, 1, nc + 1) / (float)(nc + 1) + W2_EPS) - __logf(pk_of(w, 0, nc) / w.lam_unif + W2_EPS); break;
    case 1: log_ratio = __logf(pk_of(w, 0, nc - 1) / w.lam_unif + W2_EPS) - __logf(pk_of(w, 1, nc) / (float)nc + W2_EPS); break;
    case 2: log_ratio = __logf(pk_of(w, 3, nc + 1) / (float)(nc + 1) + W2_EPS) -
                        __logf(pk_of(w, 2, nc) * dens_of(w, a.detv, a.pn0, a.pn1, a.pn2) / w.lam_data + W2_EPS); break;
    case 3: log_ratio = __logf(pk_of(w, 2, nc - 1) * dens_of(w, w.detv[r], w.pn0[r], w.pn1[r], w.pn2[r]) / w.lam_data + W2_EPS) -
                        __logf(pk_of(w, 3, nc) / (float)nc + W2_EPS); break;
    case 5: log_ratio = __logf(__fdividef(w.detv[r], aux1) + W2_EPS) - __logf(__fdividef(a.detv, aux0) + W2_EPS); break;
    default: break;  // 4, 6: symmetric; 7: set in phase 1
    }
    // ---- phase 4: Delta-energy and the accept test
    e->has_add = has_add;
    const R de = delta_lane(m, w, r, has_add, a, sx, sy);
    if (DBG && dbg_maxdiff) {
        const float diff = fabsf((float)(de - delta_lane_brute(m, w, r, has_add, a, sx, sy)));
        atomicMax(reinterpret_cast<int *>(dbg_maxdiff), __float_as_int(diff));
    }
    const float la = -(float)de / visit_temp(c, temp, it) + log_ratio;
    e->evaluated = true;
    e->accept = __logf(u01f(q1.w) + W2_EPS) < la;
}

// pair values between every staged entry and the object about to be added (what delta_staged stashes in the warp mode)
template <typename R>
__device__ __forceinline__ void fill_pairs(const ModelDev &m, const WinState<R> &w, int r, bool has_add, const Cand<R> &a, int lane, R *sx, R *sy,
                                           R *po, R *pa) {
    if (!has_add) return;
    Geo<R> ga; ga.x = a.x; ga.y = a.y; ga.hl = a.hl; ga.hw = a.hw; ga.ca = a.ca; ga.sa = a.sa;
    const R rad_a = r_sqrt_fast(a.hl * a.hl + a.hw * a.hw);
    const int n = w.n;
#pragma unroll 1
    for (int k = lane; k < n; k += 32) {
        R o = 0, al = 0;
        if (k != r && (w.flags[k] & W2_ALIVE)) {
            const int dx = w.x[k] - a.x, dy = w.y[k] - a.y, d2 = dx * dx + dy * dy;
            if (d2 <= m.ov_d2) o = pair_ov_w(m, w, k, ga, rad_a, d2, sx, sy);
            if (d2 <= m.al_d2) al = align_magnitude(geo_w(w, k), ga, m.rewarding);
        }
        po[k] = o; pa[k] = al;
    }
    __syncwarp();
}

// staged indices of the alive window objects, ascending (the order proposals pick from)
template <typename R>
__device__ __forceinline__ void rebuild_winlist(WinState<R> &w, int lane) {
    const int n = w.n;
    int cnt = 0;
#pragma unroll 1
    for (int b = 0; b < n; b += 32) {
        const int k = b + lane;
        const bool in = k < n && (w.flags[k] & (W2_ALIVE | W2_WIN)) == (W2_ALIVE | W2_WIN);
        const uint32_t bal = __ballot_sync(MPP_FULL, in);
        if (in) w.winlist[cnt + __popc(bal & ((1u << lane) - 1))] = (short)k;
        cnt += __popc(bal);
    }
    __syncwarp();
}

template <typename T> __device__ __forceinline__ T bcast(T v, int src) { return __shfl_sync(MPP_FULL, v, src); }

// Empty window (warp mode): every proposal is a birth whose candidate was drawn ahead, so the visit only has Delta-energies
// and accept tests left.  A warp evaluates 32 / G births at once, G lanes per birth (the staged objects are spread over the
// G lanes, reductions by width-G shuffles): with 8 warps and G = 8 the 32 proposals of a visit take ONE pass, and a visit
// that accepts nothing (the common case in an empty window) ends there.  Called by full warps; `it` is uniform within a
// group of G lanes (it < 0: idle group).  Returns the accept decision; *evaluated as in evaluate_proposal.
template <typename R, bool DBG>
__device__ __noinline__ bool evaluate_birth_group(const Ctx<R> &c, const WinState<R> &w, int it, int G, float temp, int lane, R *sx, R *sy, Cand<R> *out,
                                                  bool *evaluated, int *kernel_out, const int *near, int n_near, float *dbg_maxdiff, mpp_window_trace *tr) {
    const ModelDev &m = c.m;
    const int j = lane & (G - 1);
    bool live = it >= 0;
    int kernel = 0;
    Cand<R> &a = *out;
    bool full_t = false;
    if (live) {
        kernel = w.pkern[0][it];
        if (kernel == 2 && !(w.win_mass > 0.0)) live = false;
    }
    const bool drawn_t = live;
    if (live) {
        a.x = w.pc_x[0][it]; a.y = w.pc_y[0][it]; a.cls = w.pc_cls[0][it];
        a.size = w.pc_size[0][it]; a.ratio = w.pc_ratio[0][it]; a.angle = w.pc_angle[0][it];
        a.hl = w.pc_hl[0][it]; a.hw = w.pc_hw[0][it]; a.ca = w.pc_ca[0][it]; a.sa = w.pc_sa[0][it];
        a.pos = w.pc_pos[0][it]; a.dm0 = w.pc_dm0[0][it]; a.dm1 = w.pc_dm1[0][it]; a.dm2 = w.pc_dm2[0][it];
        a.detv = w.pc_detv[0][it]; a.pn0 = w.pc_pn0[0][it]; a.pn1 = w.pc_pn1[0][it]; a.pn2 = w.pc_pn2[0][it];
        if (w.cmask[((a.x >> 5) - w.cx0) * 2 + ((a.y >> 5) - w.cy0)] == 0xffffffffu) { live = false; full_t = true; }  // destination storage cell full
    }
    *kernel_out = kernel;
    // Delta-energy: the objects whose reductions the new object changes, spread over the lanes of the group
    R acc = 0, ov_add = 0, al_add = 0;
    if (live) {
        Geo<R> ga; ga.x = a.x; ga.y = a.y; ga.hl = a.hl; ga.hw = a.hw; ga.ca = a.ca; ga.sa = a.sa;
        const R rad_a = r_sqrt_fast(a.hl * a.hl + a.hw * a.hw);
        for (int i = j; i < n_near; i += G) {  // `near`: the alive staged objects within reach of the window (W2_INNER), ascending
            const int k = near[i];
            const int dx = w.x[k] - a.x, dy = w.y[k] - a.y, d2 = dx * dx + dy * dy;
            if (d2 > m.max_d2) continue;
            R ov_a = w.ov1[k], al_a = w.al1[k];
            if (d2 <= m.ov_d2) { const R o = pair_ov_w(m, w, k, ga, rad_a, d2, sx, sy); ov_a = r_max(ov_a, o); ov_add = r_max(ov_add, o); }
            if (d2 <= m.al_d2) { const R al = align_magnitude(geo_w(w, k), ga, m.rewarding); al_a = r_max(al_a, al); al_add = r_max(al_add, al); }
            acc += f_obj(m, w, k, ov_a, al_a) - w.fcur[k];
        }
    }
    for (int off = G >> 1; off > 0; off >>= 1) {
        acc += __shfl_xor_sync(MPP_FULL, acc, off);
        ov_add = r_max(ov_add, __shfl_xor_sync(MPP_FULL, ov_add, off));
        al_add = r_max(al_add, __shfl_xor_sync(MPP_FULL, al_add, off));
    }
    *evaluated = live;
    if (!live) {
        if (DBG && tr && it >= 0 && j == 0)
            trace_write<R>(tr + it, (drawn_t ? MPP_TRACE_HAS_ADD : 0u) | (full_t ? MPP_TRACE_CELL_FULL : MPP_TRACE_LEFT_WINDOW), w.pkern[0][it], 0, 0, 0u,
                           w.uid_base + (uint32_t)it, a, 0.f, 0.f, visit_temp(c, temp, it), w.pq[0][it], w.pq[1][it], w.pq[7][it]);
        return false;
    }
    Terms<R> t;
    t.pos = a.pos;
    shape_terms<R>(m, a.dm0, a.dm1, a.dm2, &t.m0, &t.m1, &t.m2);
    t.ov = ov_add; t.al = (m.rewarding ? (R)-1 : (R)1) * al_add;
    t.area = area_prior_fast<R>(m, a.hl, a.hw);
    t.ratio = r_abs((R)m.f_target_ratio - a.ratio);
    const R de = acc + combine_fast(m, t);
#ifndef MPP_TRACE
    if (DBG && dbg_maxdiff && j == 0) {
        const float diff = fabsf((float)(de - delta_lane_brute(m, w, -1, true, a, sx, sy)));
        atomicMax(reinterpret_cast<int *>(dbg_maxdiff), __float_as_int(diff));
    }
#endif
    const float fwd = kernel == 0 ? W2_DIV(pk_of(w, 0, 0), w.lam_unif) : W2_DIV(pk_of(w, 2, 0) * dens_of(w, a.detv, a.pn0, a.pn1, a.pn2), w.lam_data);
    const float bwd = pk_of(w, kernel + 1, 1);  // / (nc + 1) = 1
    const float log_ratio = __logf(bwd + W2_EPS) - __logf(fwd + W2_EPS);
    const bool accept = __logf(u01f(w.pq[7][it]) + W2_EPS) < W2_DIV(-(float)de, visit_temp(c, temp, it)) + log_ratio;
    if (DBG && tr && j == 0)
        trace_write<R>(tr + it, MPP_TRACE_HAS_ADD | MPP_TRACE_EVALUATED | (accept ? MPP_TRACE_ACCEPT : 0u), kernel, 0, 0, 0u, w.uid_base + (uint32_t)it, a,
                       (float)de, log_ratio, visit_temp(c, temp, it), w.pq[0][it], w.pq[1][it], w.pq[7][it]);
    return accept;
}

// the speculative rounds of one visit in SIMT mode (one warp)
template <typename R, bool DBG>
__device__ __forceinline__ void simt_rounds(const Ctx<R> &c, WinState<R> &w, int per_visit, float temp, uint64_t seed, uint32_t win_id, uint64_t sweep_id,
                            int lane, R *sx, R *sy, R *po, R *pa, float *dbg_maxdiff) {
    rebuild_winlist(w, lane);
    int it = 0;
    while (it < per_visit) {
        Eval<R> e;
        const int mine = it + lane;
        e.accept = false; e.evaluated = false; e.has_add = false; e.r = -1; e.kernel = 0;
        if (mine < per_visit) evaluate_lane<R, DBG>(c, w, seed, win_id, sweep_id, mine, temp, sx, sy, &e, dbg_maxdiff);
        __syncwarp();
        const uint32_t bal = __ballot_sync(MPP_FULL, e.accept);
        const int first = bal ? __ffs(bal) - 1 : 32;
        const int used = min(first + 1, min(32, per_visit - it));
        const uint32_t evb = __ballot_sync(MPP_FULL, e.evaluated && lane < used);
        if (lane == 0) { w.n_eval += __popc(evb); w.n_done += used; }
        if (first < 32) {
            Eval<R> g;  // the accepted proposal, broadcast from its lane
            g.kernel = bcast(e.kernel, first); g.r = bcast(e.r, first); g.has_add = bcast((int)e.has_add, first) != 0;
            g.evaluated = true; g.accept = true;
            g.a.x = bcast(e.a.x, first); g.a.y = bcast(e.a.y, first); g.a.cls = bcast(e.a.cls, first);
            g.a.size = bcast(e.a.size, first); g.a.ratio = bcast(e.a.ratio, first); g.a.angle = bcast(e.a.angle, first);
            g.a.hl = bcast(e.a.hl, first); g.a.hw = bcast(e.a.hw, first); g.a.ca = bcast(e.a.ca, first); g.a.sa = bcast(e.a.sa, first);
            g.a.pos = bcast(e.a.pos, first); g.a.dm0 = bcast(e.a.dm0, first); g.a.dm1 = bcast(e.a.dm1, first); g.a.dm2 = bcast(e.a.dm2, first);
            g.a.detv = bcast(e.a.detv, first); g.a.pn0 = bcast(e.a.pn0, first); g.a.pn1 = bcast(e.a.pn1, first); g.a.pn2 = bcast(e.a.pn2, first);
            if (lane == 0) {
                w.n_acc += 1;
                if (g.has_add && g.r < 0) w.n_birth += 1;
                if (!g.has_add && g.r >= 0) w.n_death += 1;
            }
            if (w.n >= W2_K && g.has_add && g.r < 0) { if (lane == 0) atomicOr(c.err, ERRF_NEIGHBOURHOOD); }
            else {
                fill_pairs(c.m, w, g.r, g.has_add, g.a, lane, sx, sy, po, pa);
                commit_proposal(c, c, w, g, it + first, lane, sx, sy, po, pa);
                rebuild_winlist(w, lane);
            }
        }
        __syncwarp();
        it += used;
    }
}

// release store of an occupancy mask (publication of a visit: records first, then masks)
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// barrier among the `n_threads` (a multiple of 32) threads of the warps that stage a visit (named barrier 1; barrier 0 is
// __syncthreads)
__device__ __forceinline__ void stage_sync(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

// One visit of window (wi, wj) of the grid shifted by (ox, oy): staging, `per_visit` proposals, publication.
// `uid_first` is the uid of the first object this visit may create.  Must be called by the whole CTA.
template <typename R, int NW, bool DBG, bool SIMT = false, bool SPLIT = false>
__device__ void window_visit(const Ctx<R> &c, const Ctx<R> &ch, WinState<R> &w, R *scratch, int wi, int wj, int ox, int oy, int per_visit, float temp,
                             uint64_t seed, uint64_t sweep_id, uint32_t uid_first, float *dbg_maxdiff) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PER_WARP = W2_SCRATCH;
    R *sx = scratch + (size_t)warp * PER_WARP, *sy = sx;  // (unused by the register-only clip; kept in the signatures)
    R *po = scratch + (size_t)warp * PER_WARP + 32, *pa = po + W2_K;
    const uint32_t win_id = (uint32_t)wi * 65536u + (uint32_t)wj;
    const ModelDev &m = c.m;
    const int px0 = 32 * wi - ox, py0 = 32 * wj - oy;
    const int x0 = max(px0, 0), x1 = min(px0 + 32, c.H), y0 = max(py0, 0), y1 = min(py0 + 32, c.W);
#ifdef MPP_TRACE
    long long t_mark[8];
    t_mark[0] = clock64();
    long long t_eval = 0, t_commit = 0;
    int n_rounds = 0;
#define MPP_MARK(i) t_mark[i] = clock64()
#else
#define MPP_MARK(i)
#endif

    // ------------------------------------------------------------------ staging
    // Two groups of warps work side by side until the barrier in front of the rounds: warps 1..npre draw the births of the
    // visit ahead (they only read the maps), warp 0 and the remaining warps stage the neighbourhood (phases A-E, synchronised
    // among themselves by named barrier 1).
    const int n_jobs = 2 * ((per_visit + 31) >> 5);                         // (kernel mixture, chunk of 32 proposals)
    const int npre = (SIMT || NW == 1) ? 0 : min(n_jobs, NW - 1);
    const bool stager = warp == 0 || warp > npre;
    const int sg = 32 * (NW - npre), sidx = warp == 0 ? lane : (warp - npre) * 32 + lane;
    if (!stager) {
        for (int job = warp - 1; job < n_jobs; job += npre)
            predraw_births(ch, w, job & 1, job >> 1, per_visit, x0, x1, y0, y1, seed, win_id, sweep_id, scratch + (size_t)warp * PER_WARP, lane);
    } else {
    // phase A (warp 0): window constants; handles / position keys of the objects within 64 px of the window
    if (warp == 0) {
        // every independent global load of the phase is issued first (one round trip instead of four dependent ones): the masks
        // of the window's own storage cells (lanes 0-3), the row masses of the window, 1 / total mass, and the masks of the
        // (at most 36) storage cells within 64 px
        const int sx0 = max(x0 - 64, 0) >> 5, sx1 = min(x1 + 63, c.H - 1) >> 5, sy0 = max(y0 - 64, 0) >> 5, sy1 = min(y1 + 63, c.W - 1) >> 5;
        const int ncw = sy1 - sy0 + 1, ncells = (sx1 - sx0 + 1) * ncw;
        int own_cell = -1;
        uint32_t own_mask = 0xffffffffu;
        if (lane < 4) {
            const int cx = (x0 >> 5) + (lane >> 1), cy = (y0 >> 5) + (lane & 1);
            if (cx < c.nx && cy < c.ny && cx <= ((x1 - 1) >> 5) && cy <= ((y1 - 1) >> 5)) {
                own_cell = cy + cx * c.ny;
                own_mask = ld_state<SPLIT>(mask_ptr<SPLIT>(c, own_cell));
            }
        }
        double rm = 0.0;
        {
            const size_t pitch = (size_t)c.W + 1;
            if (lane < x1 - x0) rm = c.rowcum[(size_t)(x0 + lane) * pitch + y1] - c.rowcum[(size_t)(x0 + lane) * pitch + y0];
        }
        const double inv_total = c.cell_cdf[c.ncell];  // [ncell] = 1 / total mass
        int cell0 = 0, cell1 = 0;
        uint32_t msk0 = 0, msk1 = 0;
        // (lane / ncw by multiply-shift: ncw <= 6 and lane + 32 < 72, exact with ceil(65536 / ncw))
        const int rcw = ncw == 1 ? 65536 : (ncw == 2 ? 32768 : (ncw == 3 ? 21846 : (ncw == 4 ? 16384 : (ncw == 5 ? 13108 : 10923))));
        const int q0r = (lane * rcw) >> 16, q1r = ((lane + 32) * rcw) >> 16;
        if (lane < ncells) { cell0 = (sy0 + lane - q0r * ncw) + (sx0 + q0r) * c.ny; msk0 = ld_state<SPLIT>(mask_ptr<SPLIT>(c, cell0)); }
        if (lane + 32 < ncells) { cell1 = (sy0 + (lane + 32) - q1r * ncw) + (sx0 + q1r) * c.ny; msk1 = ld_state<SPLIT>(mask_ptr<SPLIT>(c, cell1)); }
        // per-visit constants, spread over the lanes
        if (lane < 8) w.pkf[lane] = c.k.pf[lane];
        for (int k = lane; k < MPP_WINDOW_STATS; k += 32) w.kstat[k] = 0;
        if (lane < 4) { w.ccell[lane] = own_cell; w.cmask[lane] = own_mask; }
        w.row_mass[lane] = (float)rm;
        if (lane == 0) {
            w.x0 = x0; w.x1 = x1; w.y0 = y0; w.y1 = y1;
            w.cx0 = x0 >> 5; w.cy0 = y0 >> 5;
            w.dn = 0; w.n_acc = 0; w.n_birth = 0; w.n_death = 0; w.n_eval = 0; w.n_done = 0; w.masks_dirty = 0;
            // uids of objects born here: a function of (sweep, window, proposal index) only, so that the chain and the uids do not
            // depend on the schedule, on the proposals per visit of other calls, or on how a scene is split across GPUs.  The
            // host refuses sweep numbers whose uids would leave the 31-bit range (uid_space_ok in mpp_b200.cu).
            w.uid_base = 0x80000000u | (uint32_t)(((sweep_id * (uint64_t)((c.nx + 2) * (c.ny + 2)) + (uint64_t)(wi * (c.ny + 2) + wj)) * (uint64_t)W2_PRE) & 0x7fffffffull);
            w.pk_e0 = c.k.pk_e0; w.pk_e2 = c.k.pk_e2;
            w.dens_scale = W2_DIV((float)c.H * (float)c.W * 32768.0f, c.det_sum);
            w.lam_unif = (float)(c.k.unif_scale * (double)((x1 - x0) * (y1 - y0)));
        }
        {   // detection mass of the window
            const double tot = warp_sum(rm);
            if (lane == 0) { w.win_mass = tot; w.lam_data = (float)(c.k.intensity * tot * inv_total); }
        }
        int n = 0;
        // the 16-byte record heads, two per lane and round trip (the order in which entries are collected does not matter: phase B
        // sorts them)
        while (__any_sync(MPP_FULL, (msk0 | msk1) != 0)) {
            uint32_t h[2] = {0, 0};
            int4 head[2] = {make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
            bool have[2] = {false, false};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (msk0) { const int slot = __ffs(msk0) - 1; msk0 &= msk0 - 1; h[q] = (uint32_t)cell0 * 32u + slot; have[q] = true; }
                else if (msk1) { const int slot = __ffs(msk1) - 1; msk1 &= msk1 - 1; h[q] = (uint32_t)cell1 * 32u + slot; have[q] = true; }
                if (have[q]) head[q] = ld_state<SPLIT>(reinterpret_cast<const int4 *>(rec_ptr<SPLIT>(c, h[q])));
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const bool keep = have[q] && head[q].x >= x0 - 64 && head[q].x < x1 + 64 && head[q].y >= y0 - 64 && head[q].y < y1 + 64;
                const uint32_t kb = __ballot_sync(MPP_FULL, keep);
                if (keep) {
                    const int p = n + __popc(kb & ((1u << lane) - 1));
                    if (p < W2_K) { w.handle[p] = h[q]; w.x[p] = head[q].x * 16384 + head[q].y; w.uid[p] = (uint32_t)head[q].w; }
                }
                n += __popc(kb);
            }
        }
        if (n > W2_K) { n = W2_K; if (lane == 0) atomicOr(c.err, ERRF_NEIGHBOURHOOD); }
        if (lane == 0) { w.n = n; w.n_win = 0; }
        __syncwarp();
        MPP_MARK(1);
        // phase B (still warp 0: no barrier in between): canonical order (by pixel, then uid) so that the chain does not depend
        // on storage slot order
        // (branch-free: one 64-bit (pixel key, uid) comparison per pair, the entries read by broadcast)
#pragma unroll 1
        for (int k = lane; k < n; k += 32) {
            const unsigned long long key = ((unsigned long long)(uint32_t)w.x[k] << 32) | (unsigned long long)w.uid[k];
            int rank = 0;
#pragma unroll 4
            for (int v = 0; v < n; ++v) {
                const unsigned long long kv = ((unsigned long long)(uint32_t)w.x[v] << 32) | (unsigned long long)w.uid[v];
                rank += (int)(kv < key) + (int)((kv == key) & (v < k));
            }
            w.order[rank] = w.handle[k];
        }
    }
    if (!SIMT && NW == 1)  // a single warp does everything, one after the other
        for (int job = 0; job < n_jobs; ++job)
            predraw_births(ch, w, job & 1, job >> 1, per_visit, x0, x1, y0, y1, seed, win_id, sweep_id, scratch, lane);
    stage_sync(sg);
    MPP_MARK(2);
    const int n0 = w.n;
    // phase C: one thread per object loads its full record
    for (int p = sidx; p < n0; p += sg) {
        const uint32_t h = w.order[p];
        const Rec<R> rec = load_rec_state<SPLIT>(rec_ptr<SPLIT>(c, h));
        const bool inw = rec.x >= x0 && rec.x < x1 && rec.y >= y0 && rec.y < y1;
        const bool inner = rec.x >= x0 - 32 && rec.x < x1 + 32 && rec.y >= y0 - 32 && rec.y < y1 + 32;
        w.x[p] = rec.x; w.y[p] = rec.y; w.cls[p] = rec.cls; w.handle[p] = h; w.uid[p] = rec.uid;
        w.size[p] = rec.size; w.ratio[p] = rec.ratio; w.angle[p] = rec.angle;
        w.hl[p] = rec.hl; w.hw[p] = rec.hw; w.ca[p] = rec.ca; w.sa[p] = rec.sa; w.rad[p] = r_sqrt_fast(rec.hl * rec.hl + rec.hw * rec.hw);
        w.pos[p] = rec.e_pos; w.tm0[p] = rec.e_m[0]; w.tm1[p] = rec.e_m[1]; w.tm2[p] = rec.e_m[2];
        w.flags[p] = W2_ALIVE | (inw ? W2_WIN : 0) | (inner ? W2_INNER : 0);
        unit_form(m, w, p);
        if (inw) atomicAdd(&w.n_win, 1);
    }
    stage_sync(sg);
    MPP_MARK(3);
    if (warp == 0) rebuild_winlist(w, lane);
    // phase D: per-mark details of the window objects (deaths, translations, mark transforms): seven gathers per object;
    // phase E: partner reductions of everything a move in the window can affect; one thread per object for both
    for (int k = sidx; k < n0; k += sg) {
        if (w.flags[k] & W2_WIN) {
            float detv, pn[3], dm[3];
            gather_pixel(c, w.x[k], w.y[k], w.cls[k], &detv, pn, dm);
            w.detv[k] = detv; w.pn0[k] = pn[0]; w.pn1[k] = pn[1]; w.pn2[k] = pn[2];
            w.dm0[k] = (R)dm[0]; w.dm1[k] = (R)dm[1]; w.dm2[k] = (R)dm[2];
        }
    }
    {   // phase E, a group of 8 lanes per object
        const int gid = sidx >> 3, j = sidx & 7, n_groups = sg >> 3;
        const uint32_t gmask = 0xffu << (8 * ((threadIdx.x & 31) >> 3));
        for (int k = gid; k < n0; k += n_groups) {
            if (!(w.flags[k] & W2_INNER)) continue;
            group_top2(m, w, k, j, gmask, sx, sy);
            if (j == 0) w.fcur[k] = f_obj(m, w, k, w.ov1[k], w.al1[k]);
        }
    }
    MPP_MARK(5);
    }  // stager
    __syncthreads();
    MPP_MARK(4);

    // per-proposal trace (debug instantiation): records of this visit, see mpp_window_trace
    mpp_window_trace *tr = nullptr;
    if (DBG && c.trace && sweep_id >= c.trace_sweep0) {
        const unsigned long long first = ((sweep_id - c.trace_sweep0) * (unsigned long long)((c.nx + 2) * (c.ny + 2)) +
                                          (unsigned long long)(wi * (c.ny + 2) + wj)) * (unsigned long long)per_visit;
        if (first + (unsigned long long)per_visit <= c.trace_capacity) tr = c.trace + first;
    }
    // ------------------------------------------------------------------ speculative proposal rounds
    int it = 0;
    if (!SIMT && w.n_win == 0) {  // empty window: births only, 32 / G of them per warp at once (see evaluate_birth_group)
#ifdef MPP_TRACE
        const long long t_p = clock64();
#endif
        const int L0 = (per_visit + NW - 1) / NW;                          // proposals per warp
        const int lgL = L0 <= 1 ? 0 : (L0 <= 2 ? 1 : (L0 <= 4 ? 2 : (L0 <= 8 ? 3 : (L0 <= 16 ? 4 : 5))));  // (powers of two: shifts, no divisions)
        const int L = 1 << lgL, G = 32 >> lgL, P = min(per_visit, L * NW);
        const int mine = warp * L + (lane >> (5 - lgL));
        // only the staged objects within 32 px of the window can interact with a birth inside it: compact their indices
        // (per warp, in the pair-value stash, which is free until a commit)
        int *near = reinterpret_cast<int *>(po);
        int n_near = 0;
#pragma unroll 1
        for (int b = 0; b < w.n; b += 32) {
            const int k = b + lane;
            const bool in = k < w.n && (w.flags[k] & (W2_ALIVE | W2_INNER)) == (W2_ALIVE | W2_INNER);
            const uint32_t bal = __ballot_sync(MPP_FULL, in);
            if (in) near[n_near + __popc(bal & ((1u << lane) - 1))] = k;
            n_near += __popc(bal);
        }
        __syncwarp();
        Cand<R> a;
        bool ev = false;
        int kern = 0;
        const bool acc = evaluate_birth_group<R, DBG>(ch, w, mine < P ? mine : -1, G, temp, lane, sx, sy, &a, &ev, &kern, near, n_near, dbg_maxdiff, tr);
        __syncwarp();
        const bool head = (lane & (G - 1)) == 0;
        const uint32_t bal = __ballot_sync(MPP_FULL, acc && head), evb = __ballot_sync(MPP_FULL, ev && head);
        if (lane == 0) { w.res_accept[0][warp] = bal ? warp * L + ((__ffs(bal) - 1) >> (5 - lgL)) : 0x7fffffff; w.res_eval[warp] = (int)evb; }
        __syncthreads();
        int first = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < NW; ++q) first = min(first, w.res_accept[0][q]);
        const int used = first == 0x7fffffff ? P : first + 1;
        if (head && mine < used && ev) {
            atomicAdd(&w.kstat[kern], 1);
            if (mine == first) atomicAdd(&w.kstat[16 + kern], 1);
        }
        if (warp == 0 && lane == 0) w.kstat[34] = 1;
        if (first != 0x7fffffff && warp == (first >> lgL)) {
            const int src = (first & (L - 1)) * G;
            Eval<R> g;  // the accepted birth, broadcast from the first lane of its group
            g.kernel = bcast(kern, src); g.r = -1; g.has_add = true; g.evaluated = true; g.accept = true;
            g.a.x = bcast(a.x, src); g.a.y = bcast(a.y, src); g.a.cls = bcast(a.cls, src);
            g.a.size = bcast(a.size, src); g.a.ratio = bcast(a.ratio, src); g.a.angle = bcast(a.angle, src);
            g.a.hl = bcast(a.hl, src); g.a.hw = bcast(a.hw, src); g.a.ca = bcast(a.ca, src); g.a.sa = bcast(a.sa, src);
            g.a.pos = bcast(a.pos, src); g.a.dm0 = bcast(a.dm0, src); g.a.dm1 = bcast(a.dm1, src); g.a.dm2 = bcast(a.dm2, src);
            g.a.detv = bcast(a.detv, src); g.a.pn0 = bcast(a.pn0, src); g.a.pn1 = bcast(a.pn1, src); g.a.pn2 = bcast(a.pn2, src);
            if (w.n >= W2_K) { if (lane == 0) atomicOr(c.err, ERRF_NEIGHBOURHOOD); }
            else {
                fill_pairs(m, w, -1, true, g.a, lane, sx, sy, po, pa);
                commit_proposal<R, SPLIT>(c, ch, w, g, first, lane, sx, sy, po, pa);
            }
        }
        __syncthreads();
        it = used;
#ifdef MPP_TRACE
        t_eval += clock64() - t_p;
#endif
    }
    if (SIMT) {  // lane-per-proposal mode: the CTA is one warp
#ifdef MPP_TRACE
        const long long t_s = clock64();
#endif
        simt_rounds<R, DBG>(c, w, per_visit, temp, seed, win_id, sweep_id, lane, sx, sy, po, pa, dbg_maxdiff);
        it = per_visit;
        __syncthreads();
#ifdef MPP_TRACE
        t_eval = clock64() - t_s; n_rounds = w.n_acc + 1;
#endif
    }
    int buf = 1;  // (buffer 0 of res_accept was used by the empty-window pass)
    while (it < per_visit) {
        Eval<R> e;
        const int mine = it + warp;
#ifdef MPP_TRACE
        const long long t_a = clock64();
#endif
        if (mine < per_visit) evaluate_proposal<R, DBG>(c, ch, w, seed, win_id, sweep_id, mine, temp, lane, sx, sy, po, pa, &e, dbg_maxdiff, tr ? tr + mine : nullptr);
        else { e.accept = false; e.evaluated = false; e.has_add = false; e.r = -1; e.noop = false; e.hyp = 0; e.kernel = 0; }
        if (lane == 0) w.res_accept[buf][warp] = e.accept ? (e.noop ? 2 : 1) : 0;
        __syncthreads();
#ifdef MPP_TRACE
        const long long t_b = clock64();
#endif
        int first = NW;
#pragma unroll
        for (int q = NW - 1; q >= 0; --q) if (w.res_accept[buf][q] == 1) first = q;  // first accepted proposal that changes the state
        const int used = min(first + 1, min(NW, per_visit - it));  // proposals of the chain consumed by this round
        if (lane == 0 && warp < used && e.evaluated) {  // the visit's counters are the sums of these tallies
            atomicAdd(&w.kstat[8 * e.hyp + e.kernel], 1);
            if (e.accept && (e.noop || warp == first)) atomicAdd(&w.kstat[16 + 8 * e.hyp + e.kernel], 1);
            if (e.accept && e.noop) atomicAdd(&w.kstat[32], 1);
        }
        if (first < NW) {  // a state-changing proposal was accepted: its warp commits it, the others wait for the new state
            if (warp == first) {
                if (w.n >= W2_K && e.has_add && e.r < 0) { if (lane == 0) atomicOr(c.err, ERRF_NEIGHBOURHOOD); }
                else commit_proposal<R, SPLIT>(c, ch, w, e, mine, lane, sx, sy, po, pa);
            }
            __syncthreads();
        }
#ifdef MPP_TRACE
        t_eval += t_b - t_a; t_commit += clock64() - t_b; ++n_rounds;
#endif
        it += used;
        buf ^= 1;
    }
    __syncthreads();  // the tallies of the last round are complete before they are published
#ifdef MPP_TRACE
    if (dbg_maxdiff && threadIdx.x == 0) {
        const int slot = atomicAdd(reinterpret_cast<int *>(dbg_maxdiff + 1), 1);
        if (slot < 4000) {
            float *o = dbg_maxdiff + 8 + slot * 14;
            o[0] = (float)w.n; o[1] = (float)w.n_win; o[2] = (float)(t_mark[1] - t_mark[0]); o[3] = (float)(t_mark[2] - t_mark[1]);
            o[4] = (float)(t_mark[3] - t_mark[2]); o[5] = (float)(t_mark[4] - t_mark[3]); o[6] = (float)t_eval; o[7] = (float)t_commit;
            o[8] = (float)n_rounds; { int na = 0, ne = 0; for (int k = 0; k < 16; ++k) { ne += w.kstat[k]; na += w.kstat[16 + k]; } o[9] = (float)na; o[11] = (float)ne; } o[10] = (float)(clock64() - t_mark[0]);
            o[12] = (float)(t_mark[5] - t_mark[3]); o[13] = (float)(t_mark[4] - t_mark[5]);
        }
    }
#endif
    // publication of the state: the occupancy masks of the window's cells (the records were written by the commits).  The
    // visit's statistics follow in visit_statistics, AFTER the caller has released the visit's completion stamp: the MEMBAR of a
    // release waits for every outstanding store and atomic of the thread, and the dependents of this visit need not wait for
    // forty statistics atomics.
    if (threadIdx.x == 0 && w.masks_dirty) {
        // records before masks: a window staging these cells must never see a mask bit without its record
        if (SPLIT) {
            __threadfence_system();
            for (int q = 0; q < 4; ++q) if (w.ccell[q] >= 0) __stcg(mask_ptr<SPLIT>(c, w.ccell[q]), w.cmask[q]);
        } else {  // release stores: ordered after the records without the L1 invalidation a __threadfence() brings along
            for (int q = 0; q < 4; ++q) if (w.ccell[q] >= 0) st_release_u32(mask_ptr<SPLIT>(c, w.ccell[q]), w.cmask[q]);
        }
    }
}

// Adds the tallies of the visit that just ended to the context's counters (whole CTA; `w` must still hold the visit).
template <typename R, int NW, bool SIMT = false>
__device__ __forceinline__ void visit_statistics(const Ctx<R> &c, const WinState<R> &w, int per_visit) {
    if (!SIMT)  // (the per-kernel statistics are kept by the warp-per-proposal mode only)
        for (int k = threadIdx.x; k < MPP_WINDOW_STATS; k += 32 * NW) {
            const int v = k == 33 ? 1 : w.kstat[k];
            if (v) atomicAdd(c.kstats + k, (unsigned long long)v);
        }
    if (threadIdx.x == 0) {
        int n_done = w.n_done, n_acc = w.n_acc, n_birth = w.n_birth, n_death = w.n_death, n_eval = w.n_eval;
        if (!SIMT) {  // warp-per-proposal mode: the counters are sums of the per-kernel tallies
            n_done = per_visit; n_eval = 0; n_acc = 0;
            for (int k = 0; k < 16; ++k) { n_eval += w.kstat[k]; n_acc += w.kstat[16 + k]; }
            n_birth = w.kstat[16] + w.kstat[18] + w.kstat[24] + w.kstat[26];
            n_death = w.kstat[17] + w.kstat[19] + w.kstat[25] + w.kstat[27];
        }
        atomicAdd(c.counters + 0, (unsigned long long)n_done);
        atomicAdd(c.counters + 1, (unsigned long long)n_acc);
        atomicAdd(c.counters + 2, (unsigned long long)n_birth);
        atomicAdd(c.counters + 3, (unsigned long long)n_death);
        atomicAdd(c.counters + 4, (unsigned long long)n_eval);
        if (w.dn) atomicAdd(c.n_objects, w.dn);
    }
}

// ---- schedule 0: one launch per colour class (global barrier between colours) ------------------------
template <typename R, int NW, bool DBG, bool SIMT>
__global__ void __launch_bounds__(32 * NW) k_sweep2(const __grid_constant__ Ctx<R> c, int ci, int cj, int n_wi, int n_wj, int ox, int oy, int per_visit, float temp,
                                                   uint64_t seed, uint64_t sweep_id, uint32_t uid_base, float *dbg_maxdiff) {
    extern __shared__ __align__(16) unsigned char smem[];
    WinState<R> &w = *reinterpret_cast<WinState<R> *>(smem);
    R *scratch = reinterpret_cast<R *>(smem + ((sizeof(WinState<R>) + 15) & ~(size_t)15));
    const int a = blockIdx.x;
    if (a >= n_wi * n_wj) return;
    window_visit<R, NW, DBG, SIMT>(c, c, w, scratch, ci + 3 * (a / n_wj), cj + 3 * (a % n_wj), ox, oy, per_visit, temp, seed, sweep_id,
                             uid_base + (uint32_t)a * (uint32_t)per_visit, dbg_maxdiff);
    visit_statistics<R, NW, SIMT>(c, w, per_visit);
}

// ---- schedule 1: persistent dataflow kernel -------------------------------------------------------------
// All visits of all sweeps of one mpp_run_windows call are numbered in the order (sweep, colour, window) and claimed
// in that order from a global counter by a grid of co-resident CTAs.  A visit starts as soon as every EARLIER visit
// whose window lies within 64 px of its own has completed (same sweep: neighbours of earlier colours; previous sweep:
// every window of the previous grid within 64 px; older sweeps follow by transitivity because each grid tiles the
// image).  Dependencies always have smaller numbers, hence are already claimed by a running CTA: no deadlock.  There
// is no global barrier between colours or sweeps, so slow windows no longer hold the other SMs idle.
struct SweepPlan {
    int n_sweeps, total_tasks, dg;   // dg: pitch of the completion grids
    const int *ox, *oy;              // [n_sweeps] grid offsets
    const float *temp;               // [n_sweeps]
    const int *task_base;            // [n_sweeps + 1]
    int *done;                       // [2][dg][dg]: (local sweep index + 1) of the last completed visit, by sweep parity
    int *next_task;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Polling load of a completion stamp.  Not an acquire: ld.acquire (and every fence) is followed by a CCTL.IVALL that empties the
// SM's whole L1, i.e. the map lines (detection rows, row prefix sums, mark rows) that the proposals of BOTH resident CTAs keep
// re-reading -- once per poll of a waiting CTA.  Acquire ordering is not needed here: everything the visit reads after the wait
// that another visit may have written (occupancy masks, records) is read through L2 (__ldcg / ld.volatile), the coherence
// point, by loads that are issued after the stamp has been seen (same warp: control dependency; other warps: barrier), and the
// writer makes its stores visible at L2 before the stamp (MEMBAR of st.release).  L1 only ever holds the maps, which no
// kernel writes.
__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifndef MPP_DF_MIN_BLOCKS
#define MPP_DF_MIN_BLOCKS 2
#endif
template <typename R, int NW, bool DBG, bool SIMT>
__global__ void __launch_bounds__(32 * NW, MPP_DF_MIN_BLOCKS) k_windows_dataflow(
                                                             const __grid_constant__ Ctx<R> cpar,
                                                             const Ctx<R> *__restrict__ ctx_global, SweepPlan plan, int per_visit, uint64_t seed,
                                                             uint64_t sweep_offset, uint32_t uid_base, float *dbg_maxdiff) {
    extern __shared__ __align__(16) unsigned char smem[];
    WinState<R> &w = *reinterpret_cast<WinState<R> *>(smem);
    constexpr size_t WS = (sizeof(WinState<R>) + 15) & ~(size_t)15, SC = ((size_t)NW * W2_SCRATCH * sizeof(R) + 15) & ~(size_t)15;
    R *scratch = reinterpret_cast<R *>(smem + WS);
    // The context exists twice.  Inlined code reads it from the kernel's parameter bank (`cpar`, a __grid_constant__: its fields
    // are immediate constant operands of the instructions that use them, no load at all; +3.4 % on 2048^2, +5 % on 4096^2 over
    // reading everything from shared memory).  The out-of-line helpers get the copy in shared memory: a pointer into the
    // parameter bank would be a generic one, and without __grid_constant__ taking the address of a by-value parameter makes
    // the compiler copy it to a ~1 KB local stack frame (-5 %).
    Ctx<R> &csh = *reinterpret_cast<Ctx<R> *>(smem + WS + SC);
    const Ctx<R> &c = cpar;
    {
        const int *src = reinterpret_cast<const int *>(ctx_global);
        int *dst = reinterpret_cast<int *>(&csh);
        for (int k = threadIdx.x; k < (int)(sizeof(Ctx<R>) / 4); k += 32 * NW) dst[k] = src[k];
    }
    __shared__ int s_task;
    // the plan's small arrays in shared memory: decoding a task is a binary search, i.e. dependent loads, once per visit
    constexpr int PLAN_S = 48;
    __shared__ int s_plan[3 * PLAN_S + 1];
    __shared__ float s_temp[PLAN_S];
    if (plan.n_sweeps <= PLAN_S) {
        const int S = plan.n_sweeps;
        for (int k = threadIdx.x; k < S; k += 32 * NW) { s_plan[k] = plan.ox[k]; s_plan[PLAN_S + k] = plan.oy[k]; s_temp[k] = plan.temp[k]; }
        for (int k = threadIdx.x; k <= S; k += 32 * NW) s_plan[2 * PLAN_S + k] = plan.task_base[k];
        plan.ox = s_plan; plan.oy = s_plan + PLAN_S; plan.task_base = s_plan + 2 * PLAN_S; plan.temp = s_temp;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_task = atomicAdd(plan.next_task, 1);
    for (;;) {
        __syncthreads();  // s_task is set; the previous visit's statistics have been read out of `w`
        const int t = s_task;
        if (t >= plan.total_tasks) break;
        // decode (sweep, colour, window)
        int lo = 0, hi = plan.n_sweeps - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (plan.task_base[mid] <= t) lo = mid; else hi = mid - 1; }
        const int s = lo;
        const int ox = plan.ox[s], oy = plan.oy[s];
        const int nwx = (c.H + ox + 31) / 32, nwy = (c.W + oy + 31) / 32;
        int local = t - plan.task_base[s], col = 0, n_wj = 1;
        for (; col < 9; ++col) {
            const int ci = col / 3, cj = col % 3;
            const int a_i = ci < nwx ? (nwx - ci + 2) / 3 : 0, a_j = cj < nwy ? (nwy - cj + 2) / 3 : 0;
            if (local < a_i * a_j) { n_wj = a_j; break; }
            local -= a_i * a_j;
        }
        const int wi = col / 3 + 3 * (local / n_wj), wj = col % 3 + 3 * (local % n_wj);
        // wait for the conflicting earlier visits
        if (warp == 0) {
            int *done_cur = plan.done + (size_t)(s & 1) * plan.dg * plan.dg, *done_prev = plan.done + (size_t)((s + 1) & 1) * plan.dg * plan.dg;
            const int px0 = 32 * wi - ox, py0 = 32 * wj - oy;
            const int x0 = max(px0, 0), x1 = min(px0 + 32, c.H), y0 = max(py0, 0), y1 = min(py0 + 32, c.W);
            int pi0 = 0, pi1 = -1, pj0 = 0, pj1 = -1;
            if (s > 0) {
                const int oxp = plan.ox[s - 1], oyp = plan.oy[s - 1];
                const int nwxp = (c.H + oxp + 31) / 32, nwyp = (c.W + oyp + 31) / 32;
                pi0 = max((max(x0 - 64, 0) + oxp) >> 5, 0); pi1 = min((min(x1 - 1 + 64, c.H - 1) + oxp) >> 5, nwxp - 1);
                pj0 = max((max(y0 - 64, 0) + oyp) >> 5, 0); pj1 = min((min(y1 - 1 + 64, c.W - 1) + oyp) >> 5, nwyp - 1);
            }
            const int pw = pj1 - pj0 + 1, pn = (pi1 - pi0 + 1) * pw;
            for (;;) {
                bool ok = true;
                if (lane < 25) {  // same sweep, earlier colours, |dwi|, |dwj| <= 2
                    const int ni = wi + lane / 5 - 2, nj = wj + lane % 5 - 2;
                    if (ni >= 0 && nj >= 0 && ni < nwx && nj < nwy && (ni % 3) * 3 + nj % 3 < col)
                        ok = ld_relaxed(done_cur + ni * plan.dg + nj) >= s + 1;
                }
                for (int q = lane; q < pn; q += 32)  // every window of the previous sweep within 64 px
                    ok = ok && ld_relaxed(done_prev + (pi0 + q / pw) * plan.dg + pj0 + q % pw) >= s;
                if (__all_sync(MPP_FULL, ok)) break;
                __nanosleep(100);
            }
        }
        // no barrier here: the other warps start drawing the visit's births ahead (they only read the maps) while warp 0
        // is still waiting; window_visit's first barrier comes after warp 0 has staged the neighbourhood
        window_visit<R, NW, DBG, SIMT>(c, csh, w, scratch, wi, wj, ox, oy, per_visit, plan.temp[s], seed, sweep_offset + (uint64_t)s,
                                 uid_base + (uint32_t)t * (uint32_t)per_visit, dbg_maxdiff);
        // the next task is claimed while thread 0 publishes the masks (a global atomic with a return value is a ~1 us round trip;
        // claimed any earlier, a ready task could sit behind a long visit while other CTAs are idle)
        if (threadIdx.x == 32 % (32 * NW)) s_task = atomicAdd(plan.next_task, 1);
        __syncthreads();
        // (st.release orders every store of the CTA before the stamp -- the barrier above makes them thread 0's -- with one MEMBAR;
        // a __threadfence() in front of it was a second, sequentially consistent one plus an L1 invalidation)
        if (threadIdx.x == 0) st_release(plan.done + (size_t)(s & 1) * plan.dg * plan.dg + wi * plan.dg + wj, s + 1);
        visit_statistics<R, NW, SIMT>(c, w, per_visit);
    }
}
