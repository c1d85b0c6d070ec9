// TEST-ONLY host build of the product's per-point device functions (integrator2_b200/csrc/i2_pair.cuh),
// used to pre-check numerics of math variants on the CPU (no GPU in the build container).  It is never
// loaded by the product; the product has no CPU path.
#include "../../integrator2_b200/csrc/i2_pair.cuh"
#include "../../integrator2_b200/csrc/i2_math.cuh"
#include <cstring>
#include <algorithm>

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
    T.La = norm(T.C - T.B); T.Lb = norm(T.A - T.C); T.Lc = norm(T.B - T.A);
    T.c2 = T.Lc * T.Lb * dot(T.tc, T.tb);
    return T;
}

static d3 gp(int g, d3 A, d3 B, d3 C) {
    d3 p;
    p.x = fma(C.x, g_L[g][2], fma(B.x, g_L[g][1], A.x * g_L[g][0]));
    p.y = fma(C.y, g_L[g][2], fma(B.y, g_L[g][1], A.y * g_L[g][0]));
    p.z = fma(C.z, g_L[g][2], fma(B.z, g_L[g][1], A.z * g_L[g][0]));
    return p;
}

// grouped evaluation of one control panel (A,B,C) against T: the same sequence as grouped_eval in i2_kernels.cu, with the
// lane's own predicate in place of every warp vote
static void grouped_panel(d3 A, d3 B, d3 C, const TriJ &T, int var, double *a) {
    const bool proj = (var & 8) != 0, resid = (var & 2) == 0;
    auto raw = [&](d3 M) {
        return (var & 4) || proj ? point_terms_raw<true, true>(M, T) : ((var & 1) ? point_terms_raw<true>(M, T) : point_terms_raw<false>(M, T));
    };
    a[0] = a[1] = a[2] = a[3] = 0.0;
    int g = 0;
    while (g < g_n) {
        int gEnd = g + 1;
        while (gEnd < g_n && g_w[gEnd] == g_w[g] && gEnd - g < 6) ++gEnd;
        double pn1 = 1, pd1 = 1, pn2 = 1, pd2 = 1, pn3 = 1, pd3 = 1, zr = 1, zi = 0;
        bool flagged = false;
        for (int h = g; h < gEnd; ++h) {
            const d3 M = gp(h, A, B, C);
            double sq[3];
            PointTerms t;
            if (proj) { t = point_terms_proj(M, T, sq); flagged |= near_vertex(sq) | near_edge(t, T); }
            else { t = raw(M); flagged |= eps_screen(t); }
            flagged |= !angle_small(t);
            if (h == g) { pn1 = t.N1; pd1 = t.D1; pn2 = t.N2; pd2 = t.D2; pn3 = t.N3; pd3 = t.D3; zr = t.den; zi = t.num; continue; }
            pn1 *= t.N1; pd1 *= t.D1; pn2 *= t.N2; pd2 *= t.D2; pn3 *= t.N3; pd3 *= t.D3;
            const double zin = zi * t.num;
            zi = zi * t.den;
            zi = fma(zr, t.num, zi);
            zr = fma(zr, t.den, -zin);
        }
        const double w = g_w[g], w2 = w + w;
        double th;
        if (!flagged) {
            const int am = angle_margin(zi, zr);
            if (am >= kAngleFar) th = resid ? atan_series<4, true>(zi, zr) : atan_series<4, false>(zi, zr);
            else if (am >= kAngleTiny) th = resid ? atan_series<9, true>(zi, zr) : atan_series<9, false>(zi, zr);
            else th = resid ? atan2_fast<true>(zi, zr) : atan2_fast<false>(zi, zr);
        } else {
            pn1 = pd1 = pn2 = pd2 = pn3 = pd3 = 1.0;
            th = 0.0;
            for (int h = g; h < gEnd; ++h) {
                PointTerms u = raw(gp(h, A, B, C));
                eps_fixup(u);
                pn1 *= u.N1; pd1 *= u.D1; pn2 *= u.N2; pd2 *= u.D2; pn3 *= u.N3; pd3 *= u.D3;
                th += resid ? atan2_fast<true>(u.num, u.den) : atan2_fast<false>(u.num, u.den);
            }
        }
        const double sa = pn1 + pd1, da_ = pn1 - pd1, sb = pn2 + pd2, db_ = pn2 - pd2, sc_ = pn3 + pd3, dc_ = pn3 - pd3;
        const int rm = std::min(ratio_margin(sa, da_), std::min(ratio_margin(sb, db_), ratio_margin(sc_, dc_)));
        const double S3[3] = {sa, sb, sc_}, D3[3] = {da_, db_, dc_}, PN[3] = {pn1, pn2, pn3}, PD[3] = {pd1, pd2, pd3};
        for (int e = 0; e < 3; ++e) {
            if (rm >= kMarginVeryFar) a[e] = fma(w2, resid ? atanh_series<3, true>(S3[e], D3[e]) : atanh_series<3, false>(S3[e], D3[e]), a[e]);
            else if (rm >= kMarginFar) a[e] = fma(w2, resid ? atanh_series<6, true>(S3[e], D3[e]) : atanh_series<6, false>(S3[e], D3[e]), a[e]);
            else if (rm >= kMarginNear1) a[e] = fma(w2, resid ? atanh_series<10, true>(S3[e], D3[e]) : atanh_series<10, false>(S3[e], D3[e]), a[e]);
            else a[e] = fma(w, resid ? log_ratio<true>(PN[e], PD[e]) : log_ratio<false>(PN[e], PD[e]), a[e]);
        }
        a[3] = fma(w2, th, a[3]);
        g = gEnd;
    }
}

extern "C" {

void emu_fast_sqrt(const double *x, long long n, double *out) { for (long long k = 0; k < n; ++k) out[k] = fast_sqrt(x[k]); }
void emu_fast_rcp(const double *x, long long n, double *out) { for (long long k = 0; k < n; ++k) out[k] = fast_rcp(x[k]); }
void emu_log_ratio(const double *a, const double *b, long long n, double *out) { for (long long k = 0; k < n; ++k) out[k] = log_ratio(a[k], b[k]); }
void emu_atan2(const double *y, const double *x, long long n, double *out) { for (long long k = 0; k < n; ++k) out[k] = atan2_fast(y[k], x[k]); }

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
        } else if ((mode & 7) == 3) {
            const int var = mode >> 3;   // bit0 EDGELEN, bit1 no residual correction, bit2 DERIVE, bit3 projection form + shortcuts
            double a[4];
            grouped_panel(I.A, I.B, I.C, T, var, a);
            const double S = measures[i];
            res = vec4((S * a[0]) * T.tc + (S * a[1]) * T.ta + (S * a[2]) * T.tb, S * a[3]);
        } else {
            double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
            for (int g = 0; g < g_n; ++g) {
                const LogTheta r = mode == 1 ? theta_psi_fast<true>(gp(g, I.A, I.B, I.C), T) : theta_psi_fast<false>(gp(g, I.A, I.B, I.C), T);
                a1 = fma(g_w[g], r.t1, a1); a2 = fma(g_w[g], r.t2, a2); a3 = fma(g_w[g], r.t3, a3); a4 = fma(g_w[g], r.theta, a4);
            }
            const double S = measures[i];
            res = vec4((S * a1) * T.tc + (S * a2) * T.ta + (S * a3) * T.tb, S * a4);
        }
        out[4 * t] = res.x; out[4 * t + 1] = res.y; out[4 * t + 2] = res.z; out[4 * t + 3] = res.w;
    }
}


// regular-pair integral at a uniform refinement level, hoisted algebra (mode 1: primitives, 2: libm), children as in the kernel
static void descend_h(d3 &A, d3 &B, d3 &C, int level, int c) {
    for (int s = level - 1; s >= 0; --s) {
        const int d = (c >> (2 * s)) & 3;
        const d3 ma = 0.5 * (B + C), mb = 0.5 * (C + A), mc = 0.5 * (A + B);
        if (d == 0) { A = mc; C = ma; }
        else if (d == 1) { A = ma; B = C; C = mb; }
        else if (d == 2) { B = A; A = mb; C = mc; }
        else { A = ma; B = mb; C = mc; }
    }
}
void emu_regular_level(const double *v, const int *cells, const double *measures, const int *tasks, long long n, int mode, int level, double *out) {
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < n; ++t) {
        const int i = tasks[3 * t], j = tasks[3 * t + 1];
        const TriJ I = make_tri(v, cells, i), T = make_tri(v, cells, j);
        double Si = measures[i];
        for (int l = 0; l < level; ++l) Si *= 0.25;
        double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
        for (int c = 0; c < (1 << (2 * level)); ++c) {
            d3 A = I.A, B = I.B, C = I.C;
            descend_h(A, B, C, level, c);
            double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
            if ((mode & 7) == 3) {
                double a[4];
                grouped_panel(A, B, C, T, mode >> 3, a);
                a1 = a[0]; a2 = a[1]; a3 = a[2]; a4 = a[3];
            } else
            for (int g = 0; g < g_n; ++g) {
                const LogTheta r = mode == 1 ? theta_psi_fast<true>(gp(g, A, B, C), T) : theta_psi_fast<false>(gp(g, A, B, C), T);
                a1 = fma(g_w[g], r.t1, a1); a2 = fma(g_w[g], r.t2, a2); a3 = fma(g_w[g], r.t3, a3); a4 = fma(g_w[g], r.theta, a4);
            }
            s1 = fma(Si, a1, s1); s2 = fma(Si, a2, s2); s3 = fma(Si, a3, s3); s4 = fma(Si, a4, s4);
        }
        const d3 psi = s1 * T.tc + s2 * T.ta + s3 * T.tb;
        out[4 * t] = psi.x; out[4 * t + 1] = psi.y; out[4 * t + 2] = psi.z; out[4 * t + 3] = s4;
    }
}


// closed-form integral of the singular part (cls 0 vertex-adjacent, 1 edge-adjacent) for n tasks; out double4[n]
void emu_singular_integral(int cls, const double *v, const int *cells, const double *normals, const double *measures, const int *tasks,
                           long long n, double *out) {
    auto V = [&](int k) { return d3{v[3 * k], v[3 * k + 1], v[3 * k + 2]}; };
    auto N = [&](int c) { return d3{normals[3 * c], normals[3 * c + 1], normals[3 * c + 2]}; };
    for (long long t = 0; t < n; ++t) {
        const int i = tasks[3 * t], j = tasks[3 * t + 1];
        const tri3 ti = {cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]}, tj = {cells[3 * j], cells[3 * j + 1], cells[3 * j + 2]};
        int si, sj;
        if (cls == 0) shifts_vertex(ti, tj, si, sj); else shifts_edge(ti, tj, si, sj);
        const tri3 ri = rot_left(ti, si), rj = rot_left(tj, sj);
        bool bad = false;
        const d4 r = cls == 0 ? integral_singular_vertex(V(ri.a), V(ri.b), V(ri.c), V(rj.a), V(rj.b), V(rj.c), N(i), N(j), measures[i], &bad)
                              : integral_singular_edge(V(ri.a), V(ri.b), V(ri.c), V(rj.a), V(rj.b), V(rj.c), N(i), N(j), measures[i]);
        out[4 * t] = r.x; out[4 * t + 1] = r.y; out[4 * t + 2] = r.z; out[4 * t + 3] = r.w;
    }
}

// singular part at the centroid-ish points of triangle i (regular part integrand of the adjacent classes), strict order
void emu_singular_point(int cls, const double *v, const int *cells, const double *normals, const double *measures, const int *tasks,
                        long long n, double *out) {
    auto V = [&](int k) { return d3{v[3 * k], v[3 * k + 1], v[3 * k + 2]}; };
    auto N = [&](int c) { return d3{normals[3 * c], normals[3 * c + 1], normals[3 * c + 2]}; };
    for (long long t = 0; t < n; ++t) {
        const int i = tasks[3 * t], j = tasks[3 * t + 1];
        const tri3 ti = {cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]}, tj = {cells[3 * j], cells[3 * j + 1], cells[3 * j + 2]};
        int si, sj;
        if (cls == 0) shifts_vertex(ti, tj, si, sj); else shifts_edge(ti, tj, si, sj);
        const tri3 rj = rot_left(tj, sj);
        const d3 M = gp(1, V(ti.a), V(ti.b), V(ti.c));
        d4 r;
        if (cls == 0) { VertexSingular vs; vs.init(V(rj.a), V(rj.b), V(rj.c), N(i), N(j), measures[i]); r = vs.at(M); }
        else { EdgeSingular es; es.init(V(rj.a), V(rj.b), V(rj.c)); r = es.at(M); }
        const d4 tp = theta_psi_strict(M, V(tj.a), V(tj.b), V(tj.c));
        out[4 * t] = tp.x - r.x; out[4 * t + 1] = tp.y - r.y; out[4 * t + 2] = tp.z - r.z; out[4 * t + 3] = tp.w - r.w;
    }
}

}  // extern "C"
