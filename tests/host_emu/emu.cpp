// TEST-ONLY host build of the product's per-point device functions (integrator2_b200/csrc/i2_pair.cuh),
// used to pre-check numerics of math variants on the CPU (no GPU in the build container).  It is never
// loaded by the product; the product has no CPU path.
#include "../../integrator2_b200/csrc/i2_pair.cuh"
#include <cstring>

using namespace i2;

static double g_L[13][3], g_w[13];
static int g_n = 0;

static TriJ make_tri(const double *v, const int *cells, int j) {
    TriJ T;
    const int *t = cells + 3 * j;
    T.A = {v[3 * t[0]], v[3 * t[0] + 1], v[3 * t[0] + 2]};
    T.B = {v[3 * t[1]], v[3 * t[1] + 1], v[3 * t[1] + 2]};
    T.C = {v[3 * t[2]], v[3 * t[2] + 1], v[3 * t[2] + 2]};
    T.ta = unit(T.C - T.B); T.tb = unit(T.A - T.C); T.tc = unit(T.B - T.A);
    T.Nu = cross(T.B - T.A, T.C - T.A);
    return T;
}

static d3 gp(int g, d3 A, d3 B, d3 C) {
    d3 p;
    p.x = fma(C.x, g_L[g][2], fma(B.x, g_L[g][1], A.x * g_L[g][0]));
    p.y = fma(C.y, g_L[g][2], fma(B.y, g_L[g][1], A.y * g_L[g][0]));
    p.z = fma(C.z, g_L[g][2], fma(B.z, g_L[g][1], A.z * g_L[g][0]));
    return p;
}

extern "C" {

void emu_set_quadrature(const double *xy, const double *w, int n) {
    g_n = n;
    for (int g = 0; g < n; ++g) { g_L[g][0] = xy[2 * g]; g_L[g][1] = xy[2 * g + 1]; g_L[g][2] = 1.0 - xy[2 * g] - xy[2 * g + 1]; g_w[g] = w[g]; }
}

// regular-pair integrals at level 0: mode 0 strict, 1 fast; out double4[n]
void emu_regular(const double *v, const int *cells, const double *measures, const int *tasks, long long n, int mode, double *out) {
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < n; ++t) {
        const int i = tasks[3 * t], j = tasks[3 * t + 1];
        const TriJ I = make_tri(v, cells, i), T = make_tri(v, cells, j);
        d4 res;
        if (mode == 0) {
            d4 acc = {0, 0, 0, 0};
            for (int g = 0; g < g_n; ++g) {
                const d4 f = theta_psi_strict(gp(g, I.A, I.B, I.C), T.A, T.B, T.C);
                acc.x = fma(g_w[g], f.x, acc.x); acc.y = fma(g_w[g], f.y, acc.y); acc.z = fma(g_w[g], f.z, acc.z); acc.w = fma(g_w[g], f.w, acc.w);
            }
            res = measures[i] * acc;
        } else {
            double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
            for (int g = 0; g < g_n; ++g) {
                const LogTheta r = theta_psi_fast(gp(g, I.A, I.B, I.C), T);
                a1 = fma(g_w[g], r.t1, a1); a2 = fma(g_w[g], r.t2, a2); a3 = fma(g_w[g], r.t3, a3); a4 = fma(g_w[g], r.theta, a4);
            }
            const double S = measures[i];
            res = vec4((S * a1) * T.tc + (S * a2) * T.ta + (S * a3) * T.tb, S * a4);
        }
        out[4 * t] = res.x; out[4 * t + 1] = res.y; out[4 * t + 2] = res.z; out[4 * t + 3] = res.w;
    }
}

}  // extern "C"
