/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, float64) of the numeric core of the
 * reference's MPP hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path never does.
 *
 * PARITY UNPINNED at the shapely/GEOS boundary: the reference delegates polygon intersection to
 * shapely==1.7.1 / geos==3.8.0 (env.yml:184,50; call sites models/mpp/energies/prior_energies.py:14-18,63),
 * which is not available here and has no golden values in the reference's tests.  For convex
 * quadrilaterals the published algorithm (intersection area of two convex rings) is restated as a
 * float64 Sutherland-Hodgman clip + shoelace area.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define MAXV 16

/* base/shapes/rectangle.py:20-25 (length/width), :28-30 (poly_coord calls rect_to_poly with
 * short=length, long=width, angle+pi/2), :64-66 rotation_matrix, :69-100 rect_to_poly. */
void orc_rect_corners(double x, double y, double size, double ratio, double angle, double *out /*[8]*/)
{
    double length = (2.0 * size) / (1.0 + ratio);
    double width = ratio * length;
    double a = angle + M_PI / 2.0;
    double c = cos(a), s = sin(a);
    double lx[4] = {length / 2, length / 2, -length / 2, -length / 2};
    double ly[4] = {width / 2, -width / 2, -width / 2, width / 2};
    for (int k = 0; k < 4; ++k) {
        /* row-vector times R^T == R applied to the vector */
        out[2 * k + 0] = lx[k] * c - ly[k] * s + x;
        out[2 * k + 1] = lx[k] * s + ly[k] * c + y;
    }
}

static double signed_area(const double *p, int n)
{
    double a = 0.0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        a += p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
    }
    return 0.5 * a;
}

/* shapely Polygon.area (prior_energies.py:17,63) */
double orc_poly_area(const double *p, int n)
{
    if (n < 3) return 0.0;
    return fabs(signed_area(p, n));
}

/* shapely Polygon.intersection(other).area for convex rings (prior_energies.py:16). */
double orc_convex_intersection_area(const double *subj, int ns, const double *clip_in, int nc)
{
    double clip[2 * MAXV], bufa[2 * MAXV], bufb[2 * MAXV];
    if (ns < 3 || nc < 3) return 0.0;
    /* degenerate ring (size == 0 or ratio == 0): empty interior */
    if (signed_area(subj, ns) == 0.0 || signed_area(clip_in, nc) == 0.0) return 0.0;
    if (signed_area(clip_in, nc) < 0) {
        for (int i = 0; i < nc; ++i) {
            clip[2 * i] = clip_in[2 * (nc - 1 - i)];
            clip[2 * i + 1] = clip_in[2 * (nc - 1 - i) + 1];
        }
    } else {
        memcpy(clip, clip_in, sizeof(double) * 2 * nc);
    }
    double *in = bufa, *out = bufb;
    int n = ns;
    memcpy(in, subj, sizeof(double) * 2 * ns);
    for (int i = 0; i < nc && n > 0; ++i) {
        double ax = clip[2 * i], ay = clip[2 * i + 1];
        double bx = clip[2 * ((i + 1) % nc)], by = clip[2 * ((i + 1) % nc) + 1];
        double ex = bx - ax, ey = by - ay;
        int m = 0;
        for (int k = 0; k < n; ++k) {
            double px = in[2 * k], py = in[2 * k + 1];
            double qx = in[2 * ((k + 1) % n)], qy = in[2 * ((k + 1) % n) + 1];
            double sp = ex * (py - ay) - ey * (px - ax);
            double sq = ex * (qy - ay) - ey * (qx - ax);
            if (sp >= 0) {
                out[2 * m] = px; out[2 * m + 1] = py; ++m;
            }
            if ((sp >= 0) != (sq >= 0)) {
                double t = sp / (sp - sq);
                out[2 * m] = px + t * (qx - px); out[2 * m + 1] = py + t * (qy - py); ++m;
            }
            if (m >= MAXV - 1) break;
        }
        double *tmp = in; in = out; out = tmp;
        n = m;
    }
    if (n < 3) return 0.0;
    return fabs(signed_area(in, n));
}

/* RectangleOverlapEnergy.compute_one_interaction, prior_energies.py:13-18 */
double orc_overlap_energy(const double *r1 /*x,y,size,ratio,angle*/, const double *r2)
{
    double p1[8], p2[8];
    orc_rect_corners(r1[0], r1[1], r1[2], r1[3], r1[4], p1);
    orc_rect_corners(r2[0], r2[1], r2[2], r2[3], r2[4], p2);
    double inter = orc_convex_intersection_area(p1, 4, p2, 4);
    double a1 = orc_poly_area(p1, 4), a2 = orc_poly_area(p2, 4);
    double mn = a1 < a2 ? a1 : a2;
    return inter / (mn + 1e-6);
}

/* ShapeAlignmentEnergy.response_function, prior_energies.py:36-42 */
double orc_align_energy(double angle1, double angle2, int rewarding)
{
    return 1.0 - fabs(cos(angle1 - angle2)) - (rewarding ? 1.0 : 0.0);
}

/* Pair reductions over a whole configuration (EnergyGraph.add_point energy_graph.py:46-77 creates a pair iff
 * euclidean centre distance <= max_dist; compute_subset :108-137 reduces with max (overlap) / min (rewarding
 * alignment) / max (non rewarding), absent -> 0).  objs is [N][5] doubles.  Brute force O(N^2). */
void orc_pair_reductions(const double *objs, int n, double overlap_max_dist, double align_max_dist,
                         int rewarding, double *out_overlap, double *out_align)
{
    for (int i = 0; i < n; ++i) {
        double ov = 0.0, al = 0.0;
        int has_ov = 0, has_al = 0;
        const double *a = objs + 5 * i;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double *b = objs + 5 * j;
            double dx = a[0] - b[0], dy = a[1] - b[1];
            double d = sqrt(dx * dx + dy * dy);
            if (d <= overlap_max_dist) {
                double v = orc_overlap_energy(a, b);
                if (!has_ov || v > ov) ov = v;
                has_ov = 1;
            }
            if (d <= align_max_dist) {
                double v = orc_align_energy(a[4], b[4], rewarding);
                if (!has_al) al = v;
                else if (rewarding ? (v < al) : (v > al)) al = v;
                has_al = 1;
            }
        }
        out_overlap[i] = has_ov ? ov : 0.0;
        out_align[i] = has_al ? al : 0.0;
    }
}
