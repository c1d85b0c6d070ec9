// Per-point and per-pair functions of the J3DK evaluator (integral of the gradient of the Newtonian
// potential over a triangle pair), written for the sm_100a kernels in i2_kernels.cu.
//
// What the reference computes (all paths relative to /root/reference):
//   thetaPsi                     src/evaluators/evaluatorJ3DK.cu:266-313
//   singularPartAttached/Simple  src/evaluators/evaluatorJ3DK.cu:315-407
//   phi, q_thetaPsi(_zero,_cont) src/evaluators/evaluatorJ3DK.cu:418-601
//   integrateSingularPart*       src/evaluators/evaluatorJ3DK.cu:603-780
//   shiftsFor*                   src/evaluators/evaluatorJ3DK.cu:782-825
// How it is organised here (B200-first, not a translation):
//   * everything that depends only on the influence triangle j (unit edge tangents, edge lengths,
//     un-normalised normal) is computed ONCE per triangle by the pack kernel and read from SoA;
//   * everything that depends only on the ORIGINAL pair (i,j) of an adjacent class (rotated j,
//     direction e, delta angles, 1/sqrt(S_i), 1/|edge|) is computed once per task in a prologue
//     object and reused for all Gauss points of all refined children;
//   * the regular-pair point function works on UN-normalised vectors: no per-point divisions,
//     the triple product collapses to (M-A)·N_j, the three Psi log terms are accumulated as three
//     scalars and multiplied by the tangents once per pair.
// The file also compiles as host code (tests/host_emu) so numerics can be pre-checked on CPU.
#pragma once
#include "i2_vec.cuh"
#include "i2_math.cuh"

namespace i2 {

// ---- data of the influence triangle j as stored by the pack kernel --------------------------------
struct TriJ {
    d3 A, B, C;        // vertices in the mesh's own order
    d3 ta, tb, tc;     // unit tangents (C-B)^, (A-C)^, (B-A)^
    d3 Nu;             // (B-A) x (C-A), not normalised:  (M-A)·Nu = (M-A)x(M-B)·(M-C)
    double La, Lb, Lc; // edge lengths |C-B|, |A-C|, |B-A| (only used by the EDGELEN variant of point_terms)
    double c2;         // Lc Lb (tc·tb): only used by point_terms_proj
    int s1, s2, s3;    // near_edge thresholds hi_word(Lc, La, Lb) - kNearEdgeShift (device kernels precompute them per task)
};

// ---- regular pairs: reference operation order ("strict") ------------------------------------------
// Same sequence of FP64 operations as the reference's thetaPsi; kept as the on-device parity anchor.
I2_HD d4 theta_psi_strict(d3 M, d3 A, d3 B, d3 C) {
    d3 oa = M - A, ob = M - B, oc = M - C;
    const double la = norm(oa), lb = norm(ob), lc = norm(oc);
    oa = over(oa, la); ob = over(ob, lb); oc = over(oc, lc);
    const d3 ta = unit(C - B), tb = unit(A - C), tc = unit(B - A);
    const double rac = dot(oa, tc), rbc = dot(ob, tc), rba = dot(ob, ta);
    const double rca = dot(oc, ta), rcb = dot(oc, tb), rab = dot(oa, tb);
    // epsilon fallback of the reference as selects on the log argument (same expressions, no divergence)
    const bool f1 = fabs(rbc + 1.0) < 0.5 * EPS_PSI_THETA2, f2 = fabs(rca + 1.0) < 0.5 * EPS_PSI_THETA2, f3 = fabs(rab + 1.0) < 0.5 * EPS_PSI_THETA2;
    const double t1 = log(f1 ? lb / la : (la * (1.0 + rac)) / (lb * (1.0 + rbc)));
    const double t2 = log(f2 ? lc / lb : (lb * (1.0 + rba)) / (lc * (1.0 + rca)));
    const double t3 = log(f3 ? la / lc : (lc * (1.0 + rcb)) / (la * (1.0 + rab)));
    d4 r = vec4(t1 * tc + t2 * ta + t3 * tb);
    r.w = 2.0 * atan2(dot(cross(oa, ob), oc), 1.0 + dot(oa, ob) + dot(ob, oc) + dot(oc, oa));
    return r;
}

// ---- regular pairs: hoisted / un-normalised form ("fast") ------------------------------------------
// Per point: 3 sqrt, 3 div, 3 log, 1 atan2 and ~110 plain FP64 ops (reference: 6 sqrt, 15 div).
//   ln[ l_a(1+o_a·t_c) / (l_b(1+o_b·t_c)) ] = ln[ (l_a + d_a·t_c) / (l_b + d_b·t_c) ]
//   o_a x o_b · o_c = d_a·Nu / (l_a l_b l_c) ;  1 + Σ o·o = (l_a l_b l_c + Σ (d·d) l) / (l_a l_b l_c)
// The epsilon fallback |o_b·t_c + 1| < 0.5e-12 becomes |l_b + d_b·t_c| < 0.5e-12 l_b and is resolved
// with selects on the log argument (warp-uniform control flow, no branch).
struct LogTheta { double t1, t2, t3, theta; };

// PRIM = true : branch-free primitives of i2_math.cuh (fast_sqrt / log_ratio / atan2_fast)  -> the product path
// PRIM = false: same algebra with libdevice sqrt / log / atan2 and IEEE division             -> diagnostic variant
template <bool PRIM>
I2_HD LogTheta theta_psi_fast(d3 M, const TriJ &T) {
    const d3 da = M - T.A, db = M - T.B, dc = M - T.C;
    const double la = PRIM ? fast_sqrt(norm2(da)) : sqrt(norm2(da));
    const double lb = PRIM ? fast_sqrt(norm2(db)) : sqrt(norm2(db));
    const double lc = PRIM ? fast_sqrt(norm2(dc)) : sqrt(norm2(dc));

    const double n1 = la + dot(da, T.tc), q1 = lb + dot(db, T.tc);
    const double n2 = lb + dot(db, T.ta), q2 = lc + dot(dc, T.ta);
    const double n3 = lc + dot(dc, T.tb), q3 = la + dot(da, T.tb);
    const bool f1 = fabs(q1) < 0.5 * EPS_PSI_THETA2 * lb;
    const bool f2 = fabs(q2) < 0.5 * EPS_PSI_THETA2 * lc;
    const bool f3 = fabs(q3) < 0.5 * EPS_PSI_THETA2 * la;

    LogTheta r;
    const double num = dot(da, T.Nu);
    const double den = la * lb * lc + dot(da, db) * lc + dot(db, dc) * la + dot(dc, da) * lb;
    if (PRIM) {
        r.t1 = log_ratio(f1 ? lb : n1, f1 ? la : q1);
        r.t2 = log_ratio(f2 ? lc : n2, f2 ? lb : q2);
        r.t3 = log_ratio(f3 ? la : n3, f3 ? lc : q3);
        r.theta = 2.0 * atan2_fast(num, den);
    } else {
        r.t1 = log((f1 ? lb : n1) / (f1 ? la : q1));
        r.t2 = log((f2 ? lc : n2) / (f2 ? lb : q2));
        r.t3 = log((f3 ? la : n3) / (f3 ? lc : q3));
        r.theta = 2.0 * atan2(num, den);
    }
    return r;
}

// ---- regular pairs, grouped form: one log per edge and one atan2 per GROUP of equal-weight Gauss points ---------
// Cowper's rules repeat weights (13-point rule: 1 + 3 + 3 + 6 points), and
//     sum_{g in G} w ln(N_g/D_g) = w ln( prod N_g / prod D_g ),   sum_{g in G} w atan2(y_g,x_g) = w arg prod (x_g + i y_g),
// so a pair needs 4 x (3 logs + 1 atan2) instead of 13 x (3 logs + 1 atan2).  The argument identity holds while the
// partial angle sums stay inside (-pi, pi); every point checks |y_g| <= x_g / 2 (|angle| < pi/6, groups have <= 6
// points) and a group that fails the check anywhere in the WARP is redone point by point (warp-uniform branch).
struct PointTerms {
    double N1, D1, N2, D2, N3, D3;   // log arguments (after the epsilon selects)
    double num, den;                 // Theta_g = 2 atan2(num, den)
    double la, lb, lc;               // |M-A|, |M-B|, |M-C|
};

// EDGELEN: d_b·t_c = d_a·t_c - |AB| etc. (d_b = d_a - AB): three dot products become three subtractions.
// The log arguments are returned WITHOUT the epsilon fallback; eps_screen / eps_fixup apply it.
// DERIVE (needs EDGELEN data): B and C are not read; d_b = d_a - |AB| t_c and d_c = d_a + |CA| t_b (B = A + |AB| t_c,
// C = A - |CA| t_b), which frees 12 registers for the same number of FP64 operations.
template <bool EDGELEN = false, bool DERIVE = false>
I2_HD PointTerms point_terms_raw(d3 M, const TriJ &T) {
    const d3 da = M - T.A;
    const d3 db = DERIVE ? d3{fma(-T.Lc, T.tc.x, da.x), fma(-T.Lc, T.tc.y, da.y), fma(-T.Lc, T.tc.z, da.z)} : M - T.B;
    const d3 dc = DERIVE ? d3{fma(T.Lb, T.tb.x, da.x), fma(T.Lb, T.tb.y, da.y), fma(T.Lb, T.tb.z, da.z)} : M - T.C;
    PointTerms r;
    r.la = fast_sqrt(norm2(da)); r.lb = fast_sqrt(norm2(db)); r.lc = fast_sqrt(norm2(dc));
    const double pa = dot(da, T.tc), pb = dot(db, T.ta), pc = dot(dc, T.tb);
    r.N1 = r.la + pa; r.D1 = r.lb + (EDGELEN ? pa - T.Lc : dot(db, T.tc));
    r.N2 = r.lb + pb; r.D2 = r.lc + (EDGELEN ? pb - T.La : dot(dc, T.ta));
    r.N3 = r.lc + pc; r.D3 = r.la + (EDGELEN ? pc - T.Lb : dot(da, T.tb));
    r.num = dot(da, T.Nu);
    r.den = r.la * r.lb * r.lc + dot(da, db) * r.lc + dot(db, dc) * r.la + dot(dc, da) * r.lb;
    return r;
}

// Projection form: every length and dot product of the point follows from d_a = M - A, |d_a|^2 and the two projections
// q_b = d_a·t_b, q_c = d_a·t_c, because d_b = d_a - Lc t_c and d_c = d_a + Lb t_b:
//   |d_b|^2 = |d_a|^2 + Lc (Lc - 2 q_c)      |d_c|^2 = |d_a|^2 + Lb (Lb + 2 q_b)
//   d_a·d_b = |d_a|^2 - Lc q_c               d_c·d_a = |d_a|^2 + Lb q_b          d_b·d_c = d_c·d_a - Lc q_c - Lc Lb (t_c·t_b)
// and the log arguments are taken in the symmetric form of the same segment potential,
//   (l_a + d_a·t_c) / (l_b + d_b·t_c)  =  (l_a + l_b + Lc) / (l_a + l_b - Lc)
// (multiply out with l^2 - (d·t)^2 = squared distance to the edge line, equal for both ends), which needs no projection
// at all and has no cancellation away from the edge itself: N - D = 2 Lc exactly, so far pairs keep full relative
// accuracy where the one-sided form loses log2(distance / edge) bits.  On the line beyond either end it equals the
// reference's epsilon-fallback value l_b / l_a by itself; it differs from the reference only where D -> 0, i.e. within
// ~1e-3 edge lengths of the edge segment, which near_edge() screens (superset of the reference's fallback condition for
// points further than 2e-6 |d_b| from the vertices) and sends to the one-sided form + eps_fixup.
// 60 FP64 operations per point instead of 71.  The differences are exact in real arithmetic; in FP64 the
// squared lengths of B and C lose log2(|d_a|^2 / |d_b|^2) bits, so the caller uses this form only while |d_b|^2 and
// |d_c|^2 stay above |d_a|^2 / 16 (near_vertex(sq) reports the opposite; far pairs — the bulk — always qualify).
// sq receives (|d_a|^2, |d_b|^2, |d_c|^2).
I2_HD bool near_vertex(const double *sq) {
    return (hi_word(sq[1]) < hi_word(sq[0]) - (4 << 20)) | (hi_word(sq[2]) < hi_word(sq[0]) - (4 << 20));
}
constexpr int kNearEdgeShift = 20 << 20;   // D < 2^-20 L
I2_HD PointTerms point_terms_proj(d3 M, const TriJ &T, double *sq) {
    const d3 da = M - T.A;
    const double la2 = norm2(da);
    const double qb = dot(da, T.tb), qc = dot(da, T.tc);
    const double lb2 = fma(T.Lc, fma(-2.0, qc, T.Lc), la2);
    const double lc2 = fma(T.Lb, fma(2.0, qb, T.Lb), la2);
    sq[0] = la2; sq[1] = lb2; sq[2] = lc2;
    PointTerms r;
    r.la = fast_sqrt(la2); r.lb = fast_sqrt(lb2); r.lc = fast_sqrt(lc2);
    const double sab = r.la + r.lb, sbc = r.lb + r.lc, sca = r.lc + r.la;
    r.N1 = sab + T.Lc; r.D1 = sab - T.Lc;
    r.N2 = sbc + T.La; r.D2 = sbc - T.La;
    r.N3 = sca + T.Lb; r.D3 = sca - T.Lb;
    r.num = dot(da, T.Nu);
    const double ab = fma(-T.Lc, qc, la2), ca = fma(T.Lb, qb, la2), bc = fma(-T.Lc, qc, ca - T.c2);
    r.den = r.la * r.lb * r.lc + ab * r.lc + bc * r.la + ca * r.lb;
    return r;
}
// symmetric-form D below 2^-20 of its edge (negative rounding noise included): the point is next to the edge segment
I2_HD bool near_edge(const PointTerms &r, const TriJ &T) {
    return (hi_word(r.D1) < hi_word(T.Lc) - kNearEdgeShift) | (hi_word(r.D2) < hi_word(T.La) - kNearEdgeShift) |
           (hi_word(r.D3) < hi_word(T.Lb) - kNearEdgeShift);
}

// Integer-pipe screen for the reference's fallback test |o_b·t_c + 1| < 0.5e-12  <=>  |D| < 0.5e-12 l: positive doubles
// order like their bit patterns, so "high word of |D| < high word of l minus 40 exponent steps" (|D| < ~2^-40 l) is a
// superset of the exact condition (0.5e-12 = 2^-40.86) that costs no FP64 issue slot.  Almost never true.
I2_HD bool eps_screen(const PointTerms &r) {
    // signed compares: a (rounding-noise) negative D has a negative high word and is flagged as well
    const int k = 40 << 20;
    return (hi_word(r.D1) < hi_word(r.lb) - k) | (hi_word(r.D2) < hi_word(r.lc) - k) | (hi_word(r.D3) < hi_word(r.la) - k);
}
// exact fallback selects (slow path, taken by a whole warp when any lane passes the screen)
I2_HD void eps_fixup(PointTerms &r) {
    const bool f1 = fabs(r.D1) < 0.5 * EPS_PSI_THETA2 * r.lb;
    const bool f2 = fabs(r.D2) < 0.5 * EPS_PSI_THETA2 * r.lc;
    const bool f3 = fabs(r.D3) < 0.5 * EPS_PSI_THETA2 * r.la;
    r.N1 = f1 ? r.lb : r.N1; r.D1 = f1 ? r.la : r.D1;
    r.N2 = f2 ? r.lc : r.N2; r.D2 = f2 ? r.lb : r.D2;
    r.N3 = f3 ? r.la : r.N3; r.D3 = f3 ? r.lc : r.D3;
}
// |num| < den/2 and den > 0 on the integer pipe (conservative in the low word): the point's half solid angle is below pi/6
I2_HD bool angle_small(const PointTerms &r) {
    const int hd = hi_word(r.den);
    return (hd > 0) & ((hi_word(r.num) & 0x7fffffff) < hd - (1 << 20));
}

// reference-order variant used by host-side emulation and the diagnostics: selects applied unconditionally
template <bool EDGELEN = false>
I2_HD PointTerms point_terms(d3 M, const TriJ &T) {
    PointTerms r = point_terms_raw<EDGELEN>(M, T);
    eps_fixup(r);
    return r;
}

// ---- locating the shared vertex / edge ---------------------------------------------------------------
// positions (in i, in j) of the common vertex of a vertex-adjacent pair
I2_HD void shifts_vertex(tri3 ti, tri3 tj, int &si, int &sj) {
    si = 0; sj = 0;
    for (int p = 0; p < 3; ++p)
        for (int q = 0; q < 3; ++q)
            if (tri_at(ti, p) == tri_at(tj, q)) { si = p; sj = q; break; }
}
// positions of the vertices OPPOSITE the shared edge of an edge-adjacent pair
I2_HD void shifts_edge(tri3 ti, tri3 tj, int &si, int &sj) {
    int usedI = 0, usedJ = 0, cnt = 0;
    for (int p = 0; p < 3; ++p)
        for (int q = 0; q < 3; ++q)
            if (tri_at(ti, p) == tri_at(tj, q) && cnt < 2) { usedI |= 1 << p; usedJ |= 1 << q; ++cnt; }
    si = (usedI & 1) ? ((usedI & 2) ? 2 : 1) : 0;
    sj = (usedJ & 1) ? ((usedJ & 2) ? 2 : 1) : 0;
}

// ---- edge-adjacent pairs: singular part at a point, per-pair prologue + per-point body -----------------
struct EdgeSingular {
    d3 JB, JC, ta, tb, tc;
    double ilc;
    // tj = j rotated so that the vertex opposite the shared edge comes first
    I2_HD void init(d3 JA, d3 JB_, d3 JC_) {
        JB = JB_; JC = JC_;
        ta = unit(JA - JC_);
        tb = unit(JB_ - JA);
        const d3 e = JC_ - JB_;
        ilc = 1.0 / norm(e);
        tc = ilc * e;
    }
    I2_HD d4 at(d3 M) const {
        d3 va = M - JB, vb = M - JC;
        const double la = norm(va), lb = norm(vb);
        va = over(va, la); vb = over(vb, lb);
        const double g1 = log((lb * dot(tc, tc - vb)) / (la * dot(tc, tc - va)));
        const double g2 = log(la * dot(tb, tb + va) * ilc);
        const double g3 = log(lb * dot(ta, ta - vb) * ilc);
        d4 r = vec4(g1 * tc - g2 * tb - g3 * ta);
        r.w = 2.0 * (atan2(dot(cross(va, tb), tc), dot(tb - tc, tb + va)) -
                     atan2(dot(cross(vb, ta), tc), dot(ta - tc, ta - vb)));
        return r;
    }
};

// direction e of the intersection line of the two planes and the angles (delta_a, delta_b) it makes with
// the edges of j at the shared vertex; the two conditional flips follow the reference exactly.
struct VertexFrame {
    d3 e;
    double da, db;
    I2_HD void deltas(d3 dir, d3 ta, d3 tb, d3 nj) {
        da = atan2(dot(cross(ta, dir), nj), -dot(dir, ta));
        db = atan2(dot(cross(dir, tb), nj), dot(dir, tb));
    }
    I2_HD void init(d3 ni, d3 nj, d3 ta, d3 tb) {
        const d3 c = cross(ni, nj);
        const bool coplanar = norm2(c) < EPS_ZERO2;
        const d3 u = unit(c);                       // garbage for coplanar pairs, discarded by the select
        e = {coplanar ? tb.x : u.x, coplanar ? tb.y : u.y, coplanar ? tb.z : u.z};
        deltas(e, ta, tb, nj);
        // the two conditional flips of the reference, each evaluated by the whole warp only if some lane needs it
        const bool flip1 = (PI - fabs(da) < EPS_ZERO) || (PI - fabs(db) < EPS_ZERO);
        if (I2_WARP_ANY(flip1)) flip(flip1, ta, tb, nj);
        const bool flip2 = (da * db < 0) && (fabs(da - db) > PI);
        if (I2_WARP_ANY(flip2)) flip(flip2, ta, tb, nj);
    }
    I2_HD void flip(bool mine, d3 ta, d3 tb, d3 nj) {
        VertexFrame f;
        f.e = -1.0 * e;
        f.deltas(f.e, ta, tb, nj);
        e = {mine ? f.e.x : e.x, mine ? f.e.y : e.y, mine ? f.e.z : e.z};
        da = mine ? f.da : da;
        db = mine ? f.db : db;
    }
};

I2_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

I2_HD void sin_cos(double a, double &s, double &c) {
#if defined(__CUDA_ARCH__)
    sincos(a, &s, &c);
#else
    s = sin(a); c = cos(a);
#endif
}

// ---- vertex-adjacent pairs: singular part at a point ---------------------------------------------------
struct VertexSingular {
    d3 JA, ta, tb, e;
    double invSqrtSi;
    // (JA,JB,JC) = j rotated so that the shared vertex comes first
    I2_HD void init(d3 JA_, d3 JB, d3 JC, d3 ni, d3 nj, double Si) {
        JA = JA_;
        ta = unit(JA_ - JC);
        tb = unit(JB - JA_);
        VertexFrame fr;
        fr.init(ni, nj, ta, tb);
        e = fr.e;
        invSqrtSi = inv_sqrt(Si);
    }
    I2_HD d4 at(d3 M) const {
        d3 oc = M - JA;
        const double lc = norm(oc);
        oc = over(oc, lc);
        const double ga = log((lc * (1 + dot(ta, oc))) * invSqrtSi);
        const double gb = log((lc * (1 - dot(tb, oc))) * invSqrtSi);
        d4 r = vec4(-(ga * ta + gb * tb));
        r.w = 2.0 * (atan2(dot(cross(oc, ta), e), dot(e - oc, e - ta)) +
                     atan2(dot(cross(oc, tb), e), dot(e - oc, e + tb)));
        return r;
    }
};

// ---- closed-form integrals of the singular parts over the ORIGINAL control triangle i -------------------
struct sc { double s, c; };
I2_HD sc sc_of(double a) { sc r; sin_cos(a, r.s, r.c); return r; }
struct q2 { double theta, psi; };

I2_HD double phi_fn(sc al, sc ga, double sinXi, double cosLambda) {
    return 2.0 * atan2(sinXi * al.s * ga.s, 1.0 - al.c + ga.c + cosLambda);
}

// general (non-coplanar) q_Theta, q_Psi of an edge-adjacent pair
I2_HD q2 q_edge(sc al, sc be, sc ga, sc nu, sc xi, double cosMu, double cosLambda) {
    const double phi1 = phi_fn(al, ga, xi.s, cosLambda);
    const sc alm = {al.s, -al.c}, gam = {ga.s, -ga.c};
    const double phi2 = phi_fn(alm, gam, xi.s, cosLambda);
    const double k = 1.0 / (al.s * (1.0 - cosMu * cosMu));
    q2 r;
    r.theta = phi1 + ga.s * nu.s * k * (
                  (be.c * ga.s - xi.c * be.s * ga.c) * phi2 +
                  xi.s * be.s * (0.5 * (1.0 + cosMu) * log((1.0 + be.c) / (1.0 - nu.c)) +
                                 0.5 * (1.0 - cosMu) * log((1.0 - be.c) / (1.0 + nu.c)) +
                                 log((1.0 + cosLambda) / (1.0 - ga.c))));
    r.psi = 1.5 - k * (
                be.s * (nu.c + cosMu * cosLambda) * log(1.0 + cosLambda) +
                nu.s * (be.c + cosMu * ga.c) * log(1.0 - ga.c) +
                be.s * (1.0 - cosMu) * (nu.c - cosLambda) * log(be.s / nu.s) +
                nu.s * be.s * (be.s * ga.c - xi.c * ga.s * be.c) * log((1.0 - nu.c) / (1.0 + be.c)) +
                phi2 * xi.s * ga.s * nu.s * be.s);
    return r;
}

// coplanar limit; the reference leaves q_Theta uninitialised here (its comment says q_Theta = 0): 0 is used.
I2_HD q2 q_edge_coplanar(sc be, sc nu, double sinAlpha) {
    q2 r;
    r.theta = 0.0;
    r.psi = 1.5 - (nu.c * be.s * log(1.0 + nu.c) + nu.s * be.c * log(1.0 - be.c) + sinAlpha
                   - be.s + nu.s + be.s * nu.c * log(be.s / nu.s)) / sinAlpha;
    return r;
}

// IA..IC / JA..JC: i and j rotated so that the vertices opposite the shared edge come first
I2_HD d4 integral_singular_edge(d3 IA, d3 IB, d3 IC, d3 JA, d3 JB, d3 JC, d3 ni, d3 nj, double Si) {
    const d3 ta = unit(JA - JC), tb = unit(JB - JA), tc = unit(JC - JB);
    const double alpha = angle_between(IA - IC, IB - IC);
    const double beta = angle_between(IC - IB, IA - IB);
    const double gamma = angle_between(JC - JB, JA - JB);
    const double delta = angle_between(JB - JC, JA - JC);
    const double nuA = PI - alpha - beta;
    const double xiA = atan2(dot(cross(ni, nj), tc), dot(ni, nj));
    const sc al = sc_of(alpha), be = sc_of(beta), ga = sc_of(gamma), de = sc_of(delta), xi = sc_of(xiA);
    sc nu;
    nu.s = al.s * be.c + al.c * be.s;
    nu.c = al.s * be.s - al.c * be.c;
    const double cosSigma = -(al.c * de.c + xi.c * al.s * de.s);
    const double cosMu = -(be.c * ga.c + xi.c * be.s * ga.s);
    const double cosLambda = -(al.c * ga.c - xi.c * al.s * ga.s);
    const double cosTheta = -(be.c * de.c - xi.c * be.s * de.s);
    const double qab = nu.s * log(tan(0.5 * alpha) * tan(0.5 * nuA)) / be.s +
                       nu.s * log(tan(0.5 * beta) * tan(0.5 * nuA)) / al.s +
                       log(tan(0.5 * alpha) * tan(0.5 * beta));
    // general formula for every lane; the coplanar limit is evaluated warp-wide only when some lane needs it
    q2 qa = q_edge(al, be, ga, nu, xi, cosMu, cosLambda);
    q2 qb = q_edge(be, al, de, nu, xi, cosSigma, cosTheta);
    const bool copA = (fabs(xiA) < EPS_ZERO) && (fabs(beta - gamma) < EPS_ZERO);
    const bool copB = (fabs(xiA) < EPS_ZERO) && (fabs(alpha - delta) < EPS_ZERO);
    if (I2_WARP_ANY(copA)) { const q2 z = q_edge_coplanar(be, nu, al.s); qa.theta = copA ? z.theta : qa.theta; qa.psi = copA ? z.psi : qa.psi; }
    if (I2_WARP_ANY(copB)) { const q2 z = q_edge_coplanar(al, nu, be.s); qb.theta = copB ? z.theta : qb.theta; qb.psi = copB ? z.psi : qb.psi; }
    d4 r = vec4(Si * (qa.psi * tb + qb.psi * ta - qab * tc));
    r.w = Si * (qa.theta + qb.theta);
    return r;
}

struct VertexAngles {  // quantities shared by the two q^Theta/q^Psi evaluations (delta_a and delta_b)
    sc xi, mu, nu, psi;
    double muA, nuA, psiA, kappaA, sinKappa, logSinMu, logSinNu, sinNuPsi, sinMuPsi;
};

// q^Theta, q^Psi of a vertex-adjacent pair: four special cases tested in the reference's order, else general
I2_HD q2 q_vertex(const VertexAngles &g, double delta, sc de, double cosLambda, double cosTheta, double cosEta,
                  double cosSigma, double cosChi) {
    q2 r;
    const double logOneCosTheta = log(1.0 + cosTheta);
    const double logOneCosLambda = log(1.0 + cosLambda);
    const double Lambda1 = logOneCosLambda - logOneCosTheta + g.logSinNu - g.logSinMu;
    const double Lambda2 = log(tan(0.5 * g.nuA) * tan(0.5 * g.muA));

    const sc hd = sc_of(0.5 * delta);
    const double tanHalfDelta = hd.s / hd.c;
    const sc hmp = sc_of(0.5 * (g.muA - g.psiA)), hnp = sc_of(0.5 * (g.nuA + g.psiA));
    const double Amu = atan2(tanHalfDelta * hmp.c * g.xi.s, tanHalfDelta * hmp.c * g.xi.c + hmp.s);
    const double Anu = atan2(tanHalfDelta * hnp.s * g.xi.s, tanHalfDelta * hnp.s * g.xi.c + hnp.c);

    sc t1 = sc_of(0.5 * (g.muA - g.psiA) - 0.5 * (g.nuA + g.psiA));
    const sc t2 = sc_of(0.5 * g.kappaA);
    const double W = atan2(de.s * t2.s * g.xi.s, t2.c + de.s * t1.c * g.xi.c + de.c * t1.s);

    t1 = sc_of(delta - g.psiA);
    const double D = 1.0 / (sq(t1.s) + de.s * g.psi.s * (1.0 - g.xi.c) * (t1.c + cosSigma));
    const double G = g.psi.c * (de.s * cosSigma * g.xi.c + de.c * cosChi - sq(de.s) / g.psi.s);

    const double gent = 2.0 * (Anu * g.mu.s * g.sinNuPsi - Amu * g.nu.s * g.sinMuPsi -
                               D * g.mu.s * g.nu.s * de.s * (W * cosEta + 0.5 * g.psi.s * g.xi.s * (Lambda1 - Lambda2 * cosSigma))) /
                        (g.psi.s * g.sinKappa);
    const double gens = 0.5 * (3.0 - log(2.0)) +
                        (g.mu.s * g.nu.s / g.sinKappa) * ((logOneCosLambda - logOneCosTheta) * g.psi.c / g.psi.s +
                                                          D * (Lambda1 * de.s * cosEta / g.psi.s + Lambda2 * cosChi -
                                                               2.0 * W * de.s * g.xi.s - G * (g.logSinNu - g.logSinMu))) -
                        (g.mu.c * g.nu.s * (logOneCosLambda - g.logSinMu) + g.mu.s * g.nu.c * (logOneCosTheta - g.logSinNu)) / g.sinKappa -
                        0.5 * (g.logSinMu + g.logSinNu - log(g.sinKappa));

    const double mulPsi = sgn_dz(g.xi.s * g.psi.s);
    const double mulDelta = sgn_dz(g.xi.s * de.s);

    // The reference tests four special cases in a fixed order and falls through to the general formula; here the general
    // value is the default and each special case is evaluated by the whole warp only if some lane is in it.
    const bool planar = fabs(g.xi.s) < EPS_ZERO, psiOk = fabs(g.psi.s) > EPS_ZERO;
    const bool c1 = planar && (1.0 - fabs(cosSigma) < 0.5 * EPS_ZERO2) && psiOk;
    const bool c2 = !c1 && planar && psiOk;
    const bool c34 = !c1 && !c2 && (fabs(sin(g.psiA)) < EPS_ZERO);
    const bool c3 = c34 && (fabs(delta) > EPS_ZERO), c4 = c34 && !(fabs(delta) > EPS_ZERO);
    r.theta = gent;
    r.psi = gens;
    if (I2_WARP_ANY(c1)) {
        const double ara = arg_dz(sin(0.5 * (g.nuA + g.psiA)) * cosSigma);
        const double arb = arg_dz(cos(0.5 * (g.muA - g.psiA)) * cosSigma);
        const double th = 2.0 * mulPsi * cosSigma * g.xi.c / (g.sinKappa * g.psi.s) * (g.mu.s * g.sinNuPsi * ara - g.nu.s * g.sinMuPsi * arb);
        const double ps = 0.5 * (1.0 - log(2.0)) - 0.5 * log((1.0 - cosSigma * g.mu.c) * (1.0 + cosSigma * g.nu.c) / g.sinKappa) +
                          cosSigma * (g.mu.s - g.nu.s + 0.5 * sin(g.muA - g.nuA) * Lambda2) / g.sinKappa;
        r.theta = c1 ? th : r.theta;
        r.psi = c1 ? ps : r.psi;
    }
    if (I2_WARP_ANY(c2)) {
        const double are = arg_dz(sin(0.5 * (g.nuA + g.muA)) + sin(0.5 * (g.muA - g.nuA) - g.psiA + delta * g.xi.c));
        const double arc = arg_dz(1.0 / tan(0.5 * (g.nuA + g.psiA)) + tan(0.5 * delta) * g.xi.c);
        const double ard = arg_dz(tan(0.5 * (g.muA - g.psiA)) + tan(0.5 * delta) * g.xi.c);
        const double th = 2.0 * mulDelta * (de.s * g.mu.s * g.nu.s / cosChi * g.xi.c * are +
                                            g.mu.s * g.sinNuPsi * arc - g.nu.s * g.sinMuPsi * ard) / (g.sinKappa * g.psi.s);
        r.theta = c2 ? th : r.theta;      // psi stays the general value
    }
    if (I2_WARP_ANY(c3)) {
        const double th = 2.0 * (g.mu.s * g.nu.s * (W * de.c * g.xi.c - 0.5 * (Lambda1 - Lambda2 * cosSigma) * g.xi.s) / de.s +
                                 (Anu + mulDelta * arg_dz(sin(0.5 * (g.nuA + g.psiA)))) * g.mu.s * g.nu.c +
                                 (Amu + mulDelta * arg_dz(cos(0.5 * (g.muA - g.psiA)))) * g.mu.c * g.nu.s) / g.sinKappa;
        const double ps = 0.5 * (3.0 - log(2.0)) -
                          0.5 * (log((1.0 + cosLambda) * (1.0 + cosTheta) / g.sinKappa) - sin(g.muA - g.nuA) / g.sinKappa * Lambda1) -
                          g.mu.s * g.nu.s * ((Lambda1 * de.c - Lambda2 * g.psi.c) * g.xi.c + 2.0 * W * g.xi.s) / (g.sinKappa * de.s);
        r.theta = c3 ? th : r.theta;
        r.psi = c3 ? ps : r.psi;
    }
    if (I2_WARP_ANY(c4)) {
        const double th = 2.0 * arg_dz(g.psi.c);
        const double ps = 0.5 * (1.0 - log(2.0)) - 0.5 * log((1.0 - g.psi.c * g.mu.c) * (1.0 + g.psi.c * g.nu.c) / g.sinKappa) +
                          (0.5 * sin(g.muA - g.nuA) * Lambda1 + g.psi.c * (g.mu.s - g.nu.s)) / g.sinKappa;
        r.theta = c4 ? th : r.theta;
        r.psi = c4 ? ps : r.psi;
    }
    return r;
}

// IA..IC / JA..JC: i and j rotated so that the shared vertex comes first.
// *badOrientation is set when the pair is coplanar with opposite normals (the reference prints a warning there).
I2_HD d4 integral_singular_vertex(d3 IA, d3 IB, d3 IC, d3 JA, d3 JB, d3 JC, d3 ni, d3 nj, double Si, bool *badOrientation) {
    const d3 ta = unit(JA - JC), tb = unit(JB - JA);
    if (badOrientation) *badOrientation = (norm2(cross(ni, nj)) < EPS_ZERO2) && (dot(ni, nj) < 0);
    VertexFrame fr;
    fr.init(ni, nj, ta, tb);
    const d3 e = fr.e;
    const double xiA = atan2(dot(cross(ni, nj), e), dot(ni, nj));

    VertexAngles g;
    const sc dA = sc_of(fr.da), dB = sc_of(fr.db);
    g.xi = sc_of(xiA);
    const d3 s = IC - IB;
    g.nuA = angle_between(IA - IB, IC - IB);
    g.muA = angle_between(IB - IC, IA - IC);
    g.kappaA = angle_between(IB - IA, IC - IA);
    g.mu = sc_of(g.muA);
    g.nu = sc_of(g.nuA);
    g.logSinNu = log(g.nu.s);
    g.logSinMu = log(g.mu.s);
    g.sinKappa = sin(g.kappaA);
    g.psiA = atan2(dot(cross(e, s), ni), dot(e, s));
    g.psi = sc_of(g.psiA);
    const sc np = sc_of(g.nuA + g.psiA), mp = sc_of(g.muA - g.psiA);
    g.sinNuPsi = np.s;
    g.sinMuPsi = mp.s;

    q2 q[2];
    for (int k = 0; k < 2; ++k) {
        const sc d = k ? dB : dA;
        const double cosSigma = d.s * g.psi.s * g.xi.c + d.c * g.psi.c;
        const double cosChi = d.s * g.psi.c * g.xi.c - d.c * g.psi.s;
        const double cosEta = d.c * g.psi.s * g.xi.c - d.s * g.psi.c;
        const double cosTheta = d.s * np.s * g.xi.c + d.c * np.c;
        const double cosLambda = d.s * mp.s * g.xi.c - d.c * mp.c;
        q[k] = q_vertex(g, k ? fr.db : fr.da, d, cosLambda, cosTheta, cosEta, cosSigma, cosChi);
    }
    d4 r = vec4(Si * (q[0].psi * ta + q[1].psi * tb));
    r.w = (fabs(xiA) < EPS_ZERO) ? 0.0 : (Si * (q[0].theta - q[1].theta));
    return r;
}

// ---- final assembly J = (1/4pi) (Theta n_j + Psi x n_j), Theta wrapped into [-2 pi S_i, 2 pi S_i] for
//      vertex-adjacent pairs (src/evaluators/evaluatorJ3DK.cu:224-264) -----------------------------------------
I2_HD d3 assemble_J(d4 I, d3 nj, double Si, bool wrapTheta) {
    double theta = I.w;
    if (wrapTheta) {   // compile-time class property, not data dependent
        const double ref = TWO_PI * Si;
        const int pHi = -((int)trunc((theta - ref) / (2.0 * ref)) + 1);
        const int pLo = ((int)trunc((-ref - theta) / (2.0 * ref)) + 1);
        const int p = theta > ref ? pHi : (theta < -ref ? pLo : 0);
        theta = theta + 2.0 * p * ref;
    }
    const d3 psi = {I.x, I.y, I.z};
    return RECIPROCAL_FOUR_PI * (theta * nj + cross(psi, nj));
}

// Runge rule of the adaptive error control: true = not converged (src/evaluators/evaluator3d.cu:76-99,
// src/common/cuda_math.cuh:270-280).  2^p multiplies the COARSER value, as in the reference's code.
I2_HD bool runge_unconverged(d4 cur, d4 prev, double pow2p) {
    const d4 num = cur - prev;
    const d4 den = pow2p * prev - cur;
    d4 q;
    q.x = (fabs(num.x) < DOUBLE_MIN && fabs(den.x) < DOUBLE_MIN) ? 0.0 : num.x / den.x;
    q.y = (fabs(num.y) < DOUBLE_MIN && fabs(den.y) < DOUBLE_MIN) ? 0.0 : num.y / den.y;
    q.z = (fabs(num.z) < DOUBLE_MIN && fabs(den.z) < DOUBLE_MIN) ? 0.0 : num.z / den.z;
    q.w = (fabs(num.w) < EPS_ZERO2 && fabs(den.w) < EPS_ZERO2) ? 0.0 : num.w / den.w;
    return l1(q) > EPS_INTEGRATION;
}

}  // namespace i2
