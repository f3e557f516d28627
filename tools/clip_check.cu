// Host check of mpp_clip::quad_box_area against a float64 Sutherland-Hodgman clip (development tool):
//   g++ -O2 -x c++ -o /tmp/clip_check tools/clip_check.cu && /tmp/clip_check [percent of the full case count]
// (plain C++: mpp_clip.cuh compiles without nvcc; tests/test_clip_cpu.py runs a reduced count)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include "../mpp_cnn_rs_object_detection_b200/csrc/mpp_clip.cuh"

template <typename T>
static T sh_area(const T *qx, const T *qy, T hl, T hw) {
    std::vector<T> ax(qx, qx + 4), ay(qy, qy + 4);
    for (int pass = 0; pass < 4; ++pass) {
        const bool isx = pass < 2;
        const T sgn = (pass & 1) ? (T)-1 : (T)1, lim = isx ? hl : hw;
        std::vector<T> bx, by;
        const int n = (int)ax.size();
        if (n == 0) return (T)0;
        T px = ax[n - 1], py = ay[n - 1], pc = sgn * (isx ? px : py);
        bool pin = pc <= lim;
        for (int k = 0; k < n; ++k) {
            const T cx = ax[k], cy = ay[k], cc = sgn * (isx ? cx : cy);
            const bool cin = cc <= lim;
            if (pin != cin) {
                const T t = (lim - pc) / (cc - pc);
                T ix = px + t * (cx - px), iy = py + t * (cy - py);
                if (isx) ix = sgn * lim; else iy = sgn * lim;
                bx.push_back(ix); by.push_back(iy);
            }
            if (cin) { bx.push_back(cx); by.push_back(cy); }
            px = cx; py = cy; pc = cc; pin = cin;
        }
        ax.swap(bx); ay.swap(by);
        if (ax.size() < 3) return (T)0;
    }
    T acc = 0;
    const int n = (int)ax.size();
    for (int k = 0; k < n; ++k) { const int j = (k + n - 1) % n; acc += ax[j] * ay[k] - ax[k] * ay[j]; }
    return (acc < 0 ? -acc : acc) * (T)0.5;
}

template <typename R>
static void quad_of(double dx, double dy, double angA, double angB, double hlB, double hwB, R *qx, R *qy) {
    // same construction as overlap_energy (mpp_device.cuh), in precision R
    const R sa = (R)std::sin((R)angA), ca = (R)std::cos((R)angA), sb = (R)std::sin((R)angB), cb = (R)std::cos((R)angB);
    const R dlx = -sa * (R)dx + ca * (R)dy, dly = -ca * (R)dx - sa * (R)dy;
    const R cd = ca * cb + sa * sb, sd = sb * ca - cb * sa;
    const R lx[4] = {(R)hlB, (R)hlB, (R)-hlB, (R)-hlB}, ly[4] = {(R)hwB, (R)-hwB, (R)-hwB, (R)hwB};
    for (int k = 0; k < 4; ++k) { qx[k] = dlx + lx[k] * cd - ly[k] * sd; qy[k] = dly + lx[k] * sd + ly[k] * cd; }
}

static double urand() { return (double)rand() / ((double)RAND_MAX + 1.0); }

int main(int argc, char **argv) {
    const long scale = argc > 1 ? atol(argv[1]) : 100;  // percent of the full case count
    srand(1234);
    double worst_f = 0, worst_d = 0, worst_sh = 0, worst_thin = 0;  // worst_f: half-sides >= 1 px; worst_thin: 0.1 .. 1 px
    long n_cases = 0, n_pos = 0;
    auto run = [&](double dx, double dy, double angA, double angB, double hlA, double hwA, double hlB, double hwB, bool verbose) {
        // overlap_energy (mpp_device.cuh) clips in the frame of the thinner rectangle
        if (std::fmin((float)hlB, (float)hwB) < std::fmin((float)hlA, (float)hwA)) {
            std::swap(hlA, hlB); std::swap(hwA, hwB); std::swap(angA, angB); dx = -dx; dy = -dy;
        }
        const double thin = std::fmin(hlA, hwA);
        double qxd[4], qyd[4];
        float qxf[4], qyf[4];
        quad_of<double>(dx, dy, angA, angB, hlB, hwB, qxd, qyd);
        quad_of<float>(dx, dy, angA, angB, hlB, hwB, qxf, qyf);
        const double ref = sh_area<double>(qxd, qyd, hlA, hwA);
        const double shf = (double)sh_area<float>(qxf, qyf, (float)hlA, (float)hwA);
        const double vd = mpp_clip::quad_box_area<double>(qxd, qyd, hlA, hwA);
        const double vf = (double)mpp_clip::quad_box_area<float>(qxf, qyf, (float)hlA, (float)hwA);
        const double mnA = std::fmin(4 * hlA * hwA, 4 * hlB * hwB);
        const double ed = std::fabs(vd - ref) / mnA, ef = std::fabs(vf - ref) / mnA;
        { const double es = std::fabs(shf - ref) / mnA; if (es > worst_sh) worst_sh = es; }
        if (ed > worst_d) worst_d = ed;
        if (thin >= 1.0) { if (ef > worst_f) worst_f = ef; } else if (ef > worst_thin) worst_thin = ef;
        ++n_cases; n_pos += ref > 0;
        if (verbose || ed > 1e-9 || ef > 1e-3)
            printf("d=(%g,%g) angA=%.6f angB=%.6f A=(%g,%g) B=(%g,%g): ref %.9g  f64 %.9g  f32 %.9g   err %.2e %.2e\n", dx, dy, angA, angB, hlA, hwA,
                   hlB, hwB, ref, vd, vf, ed, ef);
    };
    const double PI = 3.14159265358979323846;
    // degenerate / structured cases
    run(0, 0, 0, 0, 4, 2, 4, 2, true);            // identical
    run(0, 4, 0, 0, 4, 2, 4, 2, true);            // shifted along an axis (collinear edges)
    run(0, 8, 0, 0, 4, 2, 4, 2, true);            // touching from outside along an edge
    run(4, 8, 0, 0, 4, 2, 4, 2, true);            // touching at a corner
    run(0, 0, 0, 0, 8, 6, 2, 1, true);            // B inside A
    run(0, 0, 0, 0, 2, 1, 8, 6, true);            // A inside B
    run(0, 0, 0, PI / 2, 4, 2, 4, 2, true);       // perpendicular cross
    run(1, 1, 0.3, 0.3, 4, 2, 4, 2, true);        // parallel, rotated
    run(3, 0, 0, PI / 4, 4, 4, 4, 4, true);       // diamond over a square
    run(0, 0, 0, PI / 4, 4, 4, 20, 20, true);     // A inside a rotated big B
    run(0, 0, 0, PI / 4, 4, 4, 2.9, 2.9, true);   // small diamond, corners poke out? (2.9*sqrt2 = 4.1)
    run(5, 5, 0.2, 1.1, 5, 2.5, 6, 1, true);
    // random: continuous angles
    for (long i = 0; i < 20000 * scale; ++i) {
        const double sA = 1 + 31 * urand(), rA = 0.1 + 0.9 * urand(), sB = 1 + 31 * urand(), rB = 0.1 + 0.9 * urand();
        const double lA = 2 * sA / (1 + rA), lB = 2 * sB / (1 + rB);
        const bool classes = (i & 1);
        const double angA = classes ? (rand() % 32) * PI / 32 : PI * urand(), angB = classes ? (rand() % 32) * PI / 32 : PI * urand();
        const int rng = 1 + rand() % 24;
        run(rand() % (2 * rng + 1) - rng, rand() % (2 * rng + 1) - rng, angA, angB, lA / 2, rA * lA / 2, lB / 2, rB * lB / 2, false);
    }
    // near-degenerate: almost parallel, almost touching
    for (long i = 0; i < 5000 * scale; ++i) {
        const double h = 2 + 6 * urand(), w = 1 + 3 * urand();
        const double eps = std::pow(10.0, -2 - 6 * urand()) * (urand() < 0.5 ? -1 : 1);
        const double ang = PI * urand();
        run(rand() % 9 - 4, rand() % 9 - 4, ang, ang + eps + (rand() % 2) * PI / 2, h, w, h, w, false);
    }
    // slivers: uniform births draw size ~ U(0, 32) and ratio ~ U(0, 1) (shape_samplers.py:136-141), half-sides down to 0.1 px here;
    // integer centre offsets and class-aligned angles as in a chain (grazing corners, collinear edges)
    for (long i = 0; i < 15000 * scale; ++i) {
        const double sA = 0.5 + 31.5 * urand(), rA = urand(), sB = 1 + 31 * urand(), rB = 0.2 + 0.8 * urand();
        const double lA = 2 * sA / (1 + rA), lB = 2 * sB / (1 + rB);
        if (rA * lA / 2 < 0.1) continue;
        const double angA = (i & 1) ? (rand() % 32) * PI / 32 : PI * urand(), angB = (i & 2) ? (rand() % 32) * PI / 32 : PI * urand();
        const int rng = 1 + rand() % 24;
        run(rand() % (2 * rng + 1) - rng, rand() % (2 * rng + 1) - rng, angA, angB, lA / 2, rA * lA / 2, lB / 2, rB * lB / 2, false);
    }
    printf("cases %ld (intersecting %ld): worst |area error| / min area  f64 %.3e  f32 %.3e (half-sides >= 1 px)  %.3e (0.1 .. 1 px)  "
           "(float32 Sutherland-Hodgman: %.3e)\n", n_cases, n_pos, worst_d, worst_f, worst_thin, worst_sh);
    return (worst_d < 1e-9 && worst_f < 1e-5 && worst_thin < 1e-4) ? 0 : 1;
}
