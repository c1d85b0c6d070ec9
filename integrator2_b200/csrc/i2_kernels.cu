// sm_100a kernels of the integrator2 hot path (see DESIGN.md for the per-kernel roofline).
//
// Replaces, behind the same host API, the reference kernels
//   kIntegrateNotNeighbors / kIntegrateRegularPartAttached / kIntegrateRegularPartSimple
//                                         (src/evaluators/evaluatorJ3DK.cu:87-205)
//   kIntegrateSingularPart{Attached,Simple}, kFinalize*Results   (same file :22-56, :224-264)
//   kSplitCell, kCountOrCreateTasks, kSumIntegrationResults, kExtractCellNeedsRefinement
//                                         (src/NumericalIntegrator3d.cu:37-195)
//   kCompareIntegrationResults, kCalculateIntegrationError, kAddReversedPairs
//                                         (src/evaluators/evaluator3d.cu:22-99)
//   kDetermineNeighborType, kCalculateCell{Normal,Center,Measure}  (src/Mesh3d.cu:23-142)
// Design differences (B200-first): no refined mesh / refined task list is ever materialised — a work item
// is (task slot, child index) and the child triangle is rebuilt on the fly by exact midpoint steps; per-task
// sums over children are warp-shuffle reductions (deterministic) instead of FP64 atomics; unconverged tasks
// are compacted on the device with warp ballots, so the adaptive loop needs no host round trip.
#include "i2_kernels.cuh"
#include "i2_pair.cuh"

#include <cstdio>
#include <cstdlib>

namespace i2 {

std::atomic<long long> g_launchCount{0};

// Gauss rule in constant memory: (L_x, L_y, L_z, w) per point, broadcast to all lanes.
__constant__ double c_gauss[MAX_GAUSS_POINTS * 4];
__constant__ int c_ngauss;
__constant__ int c_groupEnd[MAX_GAUSS_POINTS];   // 1 = last point of a run of equal weights (grouped evaluation)
__constant__ int c_ngroups;                      // the same runs as [c_groupStart[k], c_groupStart[k + 1])
__constant__ int c_groupStart[MAX_GAUSS_POINTS + 1];
__constant__ double c_pow2p;

cudaError_t upload_math_tables(cudaStream_t s) {
    return cudaMemcpyToSymbolAsync(c_mathTable, h_mathTable, sizeof(double) * I2_MATH_TABLE_SIZE, 0, cudaMemcpyHostToDevice, s);
}

cudaError_t upload_quadrature(const double *Lxyzw, int n, double pow2p, cudaStream_t s, bool *shape13) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_gauss, Lxyzw, sizeof(double) * 4 * n, 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_ngauss, &n, sizeof(int), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    // runs of equal weights (at most 6 points per run: the angle-sum argument of the grouped kernel needs <= 6)
    static int groupEnd[MAX_GAUSS_POINTS], groupStart[MAX_GAUSS_POINTS + 1], ngroups;
    int run = 0;
    ngroups = 0;
    groupStart[0] = 0;
    for (int g = 0; g < n; ++g) {
        ++run;
        const bool last = (g == n - 1) || (Lxyzw[4 * (g + 1) + 3] != Lxyzw[4 * g + 3]) || run == 6;
        groupEnd[g] = last ? 1 : 0;
        if (last) { run = 0; groupStart[++ngroups] = g + 1; }
    }
    if (shape13) *shape13 = n == 13 && ngroups == 4 && groupStart[1] == 1 && groupStart[2] == 4 && groupStart[3] == 7 && groupStart[4] == 13;
    e = cudaMemcpyToSymbolAsync(c_groupEnd, groupEnd, sizeof(int) * n, 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_groupStart, groupStart, sizeof(int) * (ngroups + 1), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_ngroups, &ngroups, sizeof(int), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_pow2p, &pow2p, sizeof(double), 0, cudaMemcpyHostToDevice, s);
}

static __device__ __forceinline__ d3 ld3(const double *__restrict__ base, int stride, int idx) {
    return {__ldg(base + idx), __ldg(base + stride + idx), __ldg(base + 2 * stride + idx)};
}
static __device__ __forceinline__ d3 ldv(const double *__restrict__ verts, int v) {
    return {__ldg(verts + 3 * v), __ldg(verts + 3 * v + 1), __ldg(verts + 3 * v + 2)};
}
static __device__ __forceinline__ tri3 ldtri(const int *__restrict__ cells, int c) {
    return {__ldg(cells + 3 * c), __ldg(cells + 3 * c + 1), __ldg(cells + 3 * c + 2)};
}

// ---------------------------------------------------------------------------------------------------------
// mesh geometry + SoA pack
// ---------------------------------------------------------------------------------------------------------
__global__ void k_geometry(const double *__restrict__ verts, const int *__restrict__ cells, int nc, double *normals,
                           double *centers, double *measures) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const tri3 t = ldtri(cells, c);
    const d3 A = ldv(verts, t.a), B = ldv(verts, t.b), C = ldv(verts, t.c);
    const d3 n = cross(B - A, C - A);
    if (normals) { const d3 u = unit(n); normals[3 * c] = u.x; normals[3 * c + 1] = u.y; normals[3 * c + 2] = u.z; }
    if (centers) { const d3 g = 0.3333333333333333 * (A + B + C); centers[3 * c] = g.x; centers[3 * c + 1] = g.y; centers[3 * c + 2] = g.z; }
    if (measures) measures[c] = norm(n) * 0.5;
}

__global__ void k_pack(const double *__restrict__ verts, const int *__restrict__ cells, const double *__restrict__ normals,
                       const double *__restrict__ measures, int nc, int stride, double *__restrict__ tri) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const tri3 t = ldtri(cells, c);
    const d3 A = ldv(verts, t.a), B = ldv(verts, t.b), C = ldv(verts, t.c);
    const d3 ta = unit(C - B), tb = unit(A - C), tc = unit(B - A);
    const d3 nu = cross(B - A, C - A);
    auto st3 = [&](int comp, d3 v) { tri[(comp + 0) * stride + c] = v.x; tri[(comp + 1) * stride + c] = v.y; tri[(comp + 2) * stride + c] = v.z; };
    st3(PK_A, A); st3(PK_B, B); st3(PK_C, C);
    st3(PK_TA, ta); st3(PK_TB, tb); st3(PK_TC, tc);
    st3(PK_NU, nu);
    st3(PK_N, {normals[3 * c], normals[3 * c + 1], normals[3 * c + 2]});
    tri[PK_S * stride + c] = measures[c];
    st3(PK_L, {norm(C - B), norm(A - C), norm(B - A)});
}

void launch_geometry(const double *verts, const int *cells, int nc, double *normals, double *centers, double *measures, cudaStream_t s) {
    if (nc > 0) { ++g_launchCount; k_geometry<<<(nc + 255) / 256, 256, 0, s>>>(verts, cells, nc, normals, centers, measures); }
}
void launch_pack(const double *verts, const int *cells, const double *normals, const double *measures, int nc, int stride, double *tri, cudaStream_t s) {
    if (nc > 0) { ++g_launchCount; k_pack<<<(nc + 255) / 256, 256, 0, s>>>(verts, cells, normals, measures, nc, stride, tri); }
}

// whole-mesh uniform split, materialised with deterministic slots (export path only)
__global__ void k_split_uniform(const double *__restrict__ vin, int nvIn, const int *__restrict__ cin, int ncIn,
                                const double *__restrict__ min, double *vout, int *cout, double *mout) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncIn) return;
    const tri3 t = ldtri(cin, c);
    const d3 A = ldv(vin, t.a), B = ldv(vin, t.b), C = ldv(vin, t.c);
    const d3 m[3] = {0.5 * (B + C), 0.5 * (C + A), 0.5 * (A + B)};
    const int v0 = nvIn + 3 * c;
    for (int k = 0; k < 3; ++k) { vout[3 * (v0 + k)] = m[k].x; vout[3 * (v0 + k) + 1] = m[k].y; vout[3 * (v0 + k) + 2] = m[k].z; }
    const int kids[4][3] = {{v0 + 2, t.b, v0}, {v0, t.c, v0 + 1}, {v0 + 1, t.a, v0 + 2}, {v0, v0 + 1, v0 + 2}};
    const double q = 0.25 * min[c];
    for (int k = 0; k < 4; ++k) {
        cout[3 * (4 * c + k)] = kids[k][0]; cout[3 * (4 * c + k) + 1] = kids[k][1]; cout[3 * (4 * c + k) + 2] = kids[k][2];
        mout[4 * c + k] = q;
    }
}
void launch_split_uniform(const double *vin, int nvIn, const int *cin, int ncIn, const double *min, double *vout, int *cout,
                          double *mout, cudaStream_t s) {
    if (ncIn > 0) { ++g_launchCount; k_split_uniform<<<(ncIn + 255) / 256, 256, 0, s>>>(vin, nvIn, cin, ncIn, min, vout, cout, mout); }
}

// ---------------------------------------------------------------------------------------------------------
// the integrate kernel: one neighbour class per instantiation
// ---------------------------------------------------------------------------------------------------------
// child `c` (base-4 digits, most significant first) of triangle (A,B,C) after `level` midpoint splits.
// Children are numbered as the reference creates them (src/NumericalIntegrator3d.cu:55-65):
//   0:(m_c,B,m_a) 1:(m_a,C,m_b) 2:(m_b,A,m_c) 3:(m_a,m_b,m_c),  m_a=(B+C)/2, m_b=(C+A)/2, m_c=(A+B)/2.
// 0.5*(x+y) is a single rounding, so the child vertices are bit-identical to the reference's refined mesh.
static __device__ __forceinline__ void descend(d3 &A, d3 &B, d3 &C, int level, int c) {
    for (int s = level - 1; s >= 0; --s) {
        const int d = (c >> (2 * s)) & 3;
        const d3 ma = 0.5 * (B + C), mb = 0.5 * (C + A), mc = 0.5 * (A + B);
        if (d == 0) { A = mc; C = ma; }
        else if (d == 1) { A = ma; B = C; C = mb; }
        else if (d == 2) { B = A; A = mb; C = mc; }
        else { A = ma; B = mb; C = mc; }
    }
}

// Gauss point g of triangle (A,B,C): r = L_x A + L_y B + L_z C, accumulated in the reference's order
// (src/NumericalIntegrator3d.cu:576-584; with nvcc's default contraction: mul, fma, fma).
static __device__ __forceinline__ d3 gauss_point(int g, d3 A, d3 B, d3 C) {
    const double lx = c_gauss[4 * g], ly = c_gauss[4 * g + 1], lz = c_gauss[4 * g + 2];
    d3 p;
    p.x = fma(C.x, lz, fma(B.x, ly, A.x * lx));
    p.y = fma(C.y, lz, fma(B.y, ly, A.y * lx));
    p.z = fma(C.z, lz, fma(B.z, ly, A.z * lx));
    return p;
}

// Lane layout shared by the integrate kernels.  A task at refinement level L has 4^L children:
//   G = min(4^L, 128) lanes cooperate on one task, each lane sums 4^L / G children;
//   G <= 32 : the G lanes sit inside one warp (32/G tasks per warp), reduction by shuffles;
//   G  > 32 : the lanes of one task span G/32 warps of the CTA (128/G tasks per CTA), reduction by shuffles + one
//             pass through shared memory in warp order.  Deep adaptive rounds have few tasks with many children, so
//             spreading a task over the whole CTA divides their serial tail by four.
// Both reductions run in a fixed order: results are bitwise reproducible.
struct LaneLayout {
    int children, G, perLane;
    __device__ explicit LaneLayout(int level) {
        children = 1 << (2 * level);
        G = children < kThreads ? children : kThreads;
        perLane = children / G;
    }
};

static __device__ __forceinline__ d4 warp_sum(d4 v, int lanes) {
    for (int off = lanes >> 1; off > 0; off >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
    }
    return v;
}

// unroll factor of the per-point loop of grouped_eval (tuning knob, -DI2_POINT_UNROLL=n; 2 measured 1 % faster than 1, 3 and 5 slower)
#ifndef I2_POINT_UNROLL
#define I2_POINT_UNROLL 2
#endif
constexpr int kPointUnroll = I2_POINT_UNROLL;

template <int CLS, int MODE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_integrate(PackedMesh pm, const int *__restrict__ tasks, const int *__restrict__ list, const int *__restrict__ countDev,
            long long countHost, int level, double *__restrict__ out) {
    __shared__ double red[kThreads / 32][4];
    const long long count = countDev ? (long long)*countDev : countHost;
    const LaneLayout lay(level);
    const int G = lay.G, perLane = lay.perLane;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ng = c_ngauss;
    const int stride = pm.stride;
    const double *__restrict__ tri = pm.tri;

    // per-lane partial sum of task slot `slot` over this lane's children (sub, sub + G, ...)
    auto lane_work = [&](int slot, int sub) -> d4 {
        d4 total = {0.0, 0.0, 0.0, 0.0};
        const int i = __ldg(tasks + 3 * (long long)slot), j = __ldg(tasks + 3 * (long long)slot + 1);
        double Si = __ldg(tri + PK_S * stride + i);
        for (int l = 0; l < level; ++l) Si *= 0.25;  // child area = parent/4 per level (exact)

        if (CLS == 2 && MODE != MATH_STRICT) {
            TriJ T;
            T.A = ld3(tri + PK_A * stride, stride, j); T.B = ld3(tri + PK_B * stride, stride, j); T.C = ld3(tri + PK_C * stride, stride, j);
            T.ta = ld3(tri + PK_TA * stride, stride, j); T.tb = ld3(tri + PK_TB * stride, stride, j); T.tc = ld3(tri + PK_TC * stride, stride, j);
            T.Nu = ld3(tri + PK_NU * stride, stride, j);
            double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;  // sums over children of S_child * sum_g w_g (t1,t2,t3,theta)
            for (int k = 0; k < perLane; ++k) {
                // control-panel vertices are re-read per child (L1-resident) to keep them out of the live register set
                d3 A = ld3(tri + PK_A * stride, stride, i), B = ld3(tri + PK_B * stride, stride, i), C = ld3(tri + PK_C * stride, stride, i);
                descend(A, B, C, level, sub + G * k);
                double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
#pragma unroll 1
                for (int g = 0; g < ng; ++g) {
                    const LogTheta v = theta_psi_fast<MODE == MATH_FAST_POINTWISE>(gauss_point(g, A, B, C), T);
                    const double w = c_gauss[4 * g + 3];
                    a1 = fma(w, v.t1, a1); a2 = fma(w, v.t2, a2); a3 = fma(w, v.t3, a3); a4 = fma(w, v.theta, a4);
                }
                s1 = fma(Si, a1, s1); s2 = fma(Si, a2, s2); s3 = fma(Si, a3, s3); s4 = fma(Si, a4, s4);
            }
            const d3 psi = s1 * T.tc + s2 * T.ta + s3 * T.tb;
            total = vec4(psi, s4);
        } else {
            const d3 JA = ld3(tri + PK_A * stride, stride, j), JB = ld3(tri + PK_B * stride, stride, j), JC = ld3(tri + PK_C * stride, stride, j);
            EdgeSingular es;
            VertexSingular vs;
            if (CLS == 1) {
                int si, sj;
                shifts_edge(ldtri(pm.cells, i), ldtri(pm.cells, j), si, sj);
                const d3 RA = sj == 0 ? JA : (sj == 1 ? JB : JC), RB = sj == 0 ? JB : (sj == 1 ? JC : JA), RC = sj == 0 ? JC : (sj == 1 ? JA : JB);
                es.init(RA, RB, RC);
            }
            if (CLS == 0) {
                int si, sj;
                shifts_vertex(ldtri(pm.cells, i), ldtri(pm.cells, j), si, sj);
                const d3 RA = sj == 0 ? JA : (sj == 1 ? JB : JC), RB = sj == 0 ? JB : (sj == 1 ? JC : JA), RC = sj == 0 ? JC : (sj == 1 ? JA : JB);
                vs.init(RA, RB, RC, ld3(tri + PK_N * stride, stride, i), ld3(tri + PK_N * stride, stride, j), __ldg(tri + PK_S * stride + i));
            }
            for (int k = 0; k < perLane; ++k) {
                d3 A = ld3(tri + PK_A * stride, stride, i), B = ld3(tri + PK_B * stride, stride, i), C = ld3(tri + PK_C * stride, stride, i);
                descend(A, B, C, level, sub + G * k);
                d4 acc = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
                for (int g = 0; g < ng; ++g) {
                    const d3 M = gauss_point(g, A, B, C);
                    d4 f = theta_psi_strict(M, JA, JB, JC);
                    if (CLS == 1) f = f - es.at(M);
                    if (CLS == 0) f = f - vs.at(M);
                    const double w = c_gauss[4 * g + 3];
                    acc.x = fma(w, f.x, acc.x); acc.y = fma(w, f.y, acc.y); acc.z = fma(w, f.z, acc.z); acc.w = fma(w, f.w, acc.w);
                }
                total = total + Si * acc;
            }
        }
        return total;
    };
    auto store = [&](int slot, d4 total) {
        double2 *o = reinterpret_cast<double2 *>(out + 4 * (long long)slot);
        o[0] = make_double2(total.x, total.y);
        o[1] = make_double2(total.z, total.w);
    };

    if (G <= 32) {
        const int groupsPerWarp = 32 / G, sub = lane & (G - 1);
        const long long warpId = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
        const long long warpStride = ((long long)gridDim.x * kThreads) >> 5;
        for (long long base = warpId * groupsPerWarp; base < count; base += warpStride * groupsPerWarp) {
            long long r = base + lane / G;
            const bool active = r < count;
            if (!active) r = count - 1;   // tail lanes recompute the last task (no write): warp votes in the prologues stay full-mask
            const int slot = list ? __ldg(list + r) : (int)r;
            const d4 total = warp_sum(lane_work(slot, sub), G);
            if (active && sub == 0) store(slot, total);
        }
    } else {
        const int tasksPerCTA = kThreads / G, sub = threadIdx.x & (G - 1), tIdx = threadIdx.x / G, warpsPerTask = G / 32;
        for (long long base = (long long)blockIdx.x * tasksPerCTA; base < count; base += (long long)gridDim.x * tasksPerCTA) {
            long long r = base + tIdx;
            const bool active = r < count;
            if (!active) r = count - 1;
            const int slot = list ? __ldg(list + r) : (int)r;
            const d4 part = warp_sum(lane_work(slot, sub), 32);
            if (lane == 0) { red[warp][0] = part.x; red[warp][1] = part.y; red[warp][2] = part.z; red[warp][3] = part.w; }
            __syncthreads();
            if (active && sub == 0) {
                d4 total = {0.0, 0.0, 0.0, 0.0};
                for (int w = 0; w < warpsPerTask; ++w) { total.x += red[warp + w][0]; total.y += red[warp + w][1]; total.z += red[warp + w][2]; total.w += red[warp + w][3]; }
                store(slot, total);
            }
            __syncthreads();
        }
    }
}

// Sticky "this group needs the careful form" flag of grouped_eval, raised on the integer pipe: one chained predicate
//   one-sided form : D_k below 2^-40 of its length (eps_screen)
//   projection form: symmetric D_k below 2^-20 of its edge (near_edge) | |d_b|^2 or |d_c|^2 below |d_a|^2 / 16 (near_vertex)
//   both           : half solid angle not provably below pi/6 (!angle_small; a non-positive den fails the compare by itself)
// and one select.  Written in PTX because the compiler otherwise emits one select per condition.
template <bool PROJ>
static __device__ __forceinline__ int raise_flag(int flagged, const PointTerms &t, const TriJ &T, const double *sq) {
    const int k = 40 << 20;
    int e1 = hi_word(t.lb) - k, e2 = hi_word(t.lc) - k, e3 = hi_word(t.la) - k;
    const int hn = hi_word(t.num) & 0x7fffffff, hd = hi_word(t.den) - (1 << 20);
    if (PROJ) {
        const int nv = hi_word(sq[0]) - (4 << 20);
        e1 = T.s1; e2 = T.s2; e3 = T.s3;
        asm("{\n\t.reg .pred p;\n\t"
            "setp.lt.s32 p, %1, %2;\n\t"
            "setp.lt.or.s32 p, %3, %4, p;\n\t"
            "setp.lt.or.s32 p, %5, %6, p;\n\t"
            "setp.ge.or.s32 p, %7, %8, p;\n\t"
            "setp.lt.or.s32 p, %9, %11, p;\n\t"
            "setp.lt.or.s32 p, %10, %11, p;\n\t"
            "selp.s32 %0, 1, %0, p;\n\t}"
            : "+r"(flagged)
            : "r"(hi_word(t.D1)), "r"(e1), "r"(hi_word(t.D2)), "r"(e2), "r"(hi_word(t.D3)), "r"(e3), "r"(hn), "r"(hd),
              "r"(hi_word(sq[1])), "r"(hi_word(sq[2])), "r"(nv));
    } else {
        asm("{\n\t.reg .pred p;\n\t"
            "setp.lt.s32 p, %1, %2;\n\t"
            "setp.lt.or.s32 p, %3, %4, p;\n\t"
            "setp.lt.or.s32 p, %5, %6, p;\n\t"
            "setp.ge.or.s32 p, %7, %8, p;\n\t"
            "selp.s32 %0, 1, %0, p;\n\t}"
            : "+r"(flagged)
            : "r"(hi_word(t.D1)), "r"(e1), "r"(hi_word(t.D2)), "r"(e2), "r"(hi_word(t.D3)), "r"(e3), "r"(hn), "r"(hd));
    }
    return flagged;
}

// Grouped evaluation of one (child) control panel against triangle T: on return a1..a3 = sum_g w_g ln(N/D) per edge,
// a4 = sum_g w_g Theta_g.  myM points at this thread's staged Gauss points ([point][component], stride kThreads).
// Must be called by all 32 lanes of a warp.  The point loop carries no vote and no branch: every rare condition (a log
// argument below the reference's epsilon, a large solid angle, a Gauss point next to a vertex of T in the projection form)
// only raises a sticky flag, and ONE __all_sync per group of equal weights decides whether the warp redoes the group point
// by point in the careful form.
// MS = distance (in doubles) between consecutive components of the staged points (kThreads when every thread has its own set)
// vmask = the lanes that decide together (the whole warp, or the 16 lanes of one task in the list-driven round 2 of the
// adaptive queue, where the two tasks that share a warp depend on which other tasks are still unconverged: a task's
// result must not depend on its warp-mate, or it would change with the partition of the list over GPUs)
// FIX = 13: the rule has the group structure of Cowper's 13-point rule (runs of equal weights 1 + 3 + 3 + 6, checked on the host when
// the rule is uploaded): the four groups and their points become straight-line code — no loop counters, no constant-bank loads of
// the group table, the weights as immediate constant operands — and the compiler may overlap the loads / seeds of one point with
// the arithmetic of the previous one.  Same operations in the same order: results are bit-identical to the generic loop.
template <bool EDGELEN, bool RESID, bool DERIVE = false, bool PROJ = false, int MS = kThreads, int UNR = kPointUnroll, int FIX = 0>
static __device__ __forceinline__ void grouped_eval(const double *myM, int ng, const TriJ &T, double &a1, double &a2, double &a3, double &a4,
                                                    const unsigned vmask = 0xffffffffu) {
    a1 = 0.0; a2 = 0.0; a3 = 0.0; a4 = 0.0;
    auto do_group = [&](const int gStart, const int gEnd) {
        const double *pM = myM + 3 * MS * gStart;
        double sq[3];
        auto point = [&]() -> PointTerms {
            const d3 M = {pM[0], pM[MS], pM[2 * MS]};
            pM += 3 * MS;
            return PROJ ? point_terms_proj(M, T, sq) : point_terms_raw<EDGELEN, DERIVE>(M, T);
        };
        // first point of the group: the running products start from its terms (no multiplication by one)
        PointTerms t = point();
        double pn1 = t.N1, pd1 = t.D1, pn2 = t.N2, pd2 = t.D2, pn3 = t.N3, pd3 = t.D3, zr = t.den, zi = t.num;
        int flagged = raise_flag<PROJ>(0, t, T, sq);
        auto next_point = [&]() {
            t = point();
            flagged = raise_flag<PROJ>(flagged, t, T, sq);
            pn1 *= t.N1; pd1 *= t.D1; pn2 *= t.N2; pd2 *= t.D2; pn3 *= t.N3; pd3 *= t.D3;
            const double zin = zi * t.num;   // complex product, updated in place
            zi = zi * t.den;
            zi = fma(zr, t.num, zi);
            zr = fma(zr, t.den, -zin);
        };
        if (FIX) {
#pragma unroll
            for (int p = gStart + 1; p < gEnd; ++p) next_point();
        } else {
#pragma unroll UNR
            for (int p = gStart + 1; p < gEnd; ++p) next_point();
        }
        const double w = c_gauss[4 * gStart + 3], w2 = w + w;
        double th;
        if (__all_sync(vmask, !flagged)) {
            const int am = __reduce_min_sync(vmask, angle_margin(zi, zr));   // far-field tiers, taken by the whole warp
            if (am >= kAngleFar) th = atan_series<4, RESID>(zi, zr);
            else if (am >= kAngleTiny) th = atan_series<9, RESID>(zi, zr);
            else th = atan2_fast<RESID>(zi, zr);
        } else {
            // careful form for the whole warp: epsilon fallback applied, angles added one by one
            pn1 = pd1 = pn2 = pd2 = pn3 = pd3 = 1.0;
            th = 0.0;
#pragma unroll 1
            for (int h = gStart; h < gEnd; ++h) {
                const d3 Mh = {myM[(3 * h + 0) * MS], myM[(3 * h + 1) * MS], myM[(3 * h + 2) * MS]};
                PointTerms u = point_terms_raw<EDGELEN, DERIVE>(Mh, T);
                eps_fixup(u);
                pn1 *= u.N1; pd1 *= u.D1; pn2 *= u.N2; pd2 *= u.D2; pn3 *= u.N3; pd3 *= u.D3;
                th += atan2_fast<RESID>(u.num, u.den);
            }
        }
        // far-field tiers, taken by the whole warp: all three ratios close enough to 1 -> no mantissa surgery, shorter series
        const double sa = pn1 + pd1, da_ = pn1 - pd1, sb = pn2 + pd2, db_ = pn2 - pd2, sc_ = pn3 + pd3, dc_ = pn3 - pd3;
        const int rm = __reduce_min_sync(vmask, min(ratio_margin(sa, da_), min(ratio_margin(sb, db_), ratio_margin(sc_, dc_))));
        if (rm >= kMarginVeryFar) {
            a1 = fma(w2, atanh_series<3, RESID>(sa, da_), a1);
            a2 = fma(w2, atanh_series<3, RESID>(sb, db_), a2);
            a3 = fma(w2, atanh_series<3, RESID>(sc_, dc_), a3);
        } else if (rm >= kMarginFar) {
            a1 = fma(w2, atanh_series<6, RESID>(sa, da_), a1);
            a2 = fma(w2, atanh_series<6, RESID>(sb, db_), a2);
            a3 = fma(w2, atanh_series<6, RESID>(sc_, dc_), a3);
        } else if (rm >= kMarginNear1) {
            a1 = fma(w2, atanh_series<10, RESID>(sa, da_), a1);
            a2 = fma(w2, atanh_series<10, RESID>(sb, db_), a2);
            a3 = fma(w2, atanh_series<10, RESID>(sc_, dc_), a3);
        } else {
            a1 = fma(w, log_ratio<RESID>(pn1, pd1), a1);
            a2 = fma(w, log_ratio<RESID>(pn2, pd2), a2);
            a3 = fma(w, log_ratio<RESID>(pn3, pd3), a3);
        }
        a4 = fma(w2, th, a4);
    };
    if (FIX == 13) {
        do_group(0, 1); do_group(1, 4); do_group(4, 7); do_group(7, 13);
    } else {
        const int ngroups = c_ngroups;
        int g = 0;
#pragma unroll 1
        for (int grp = 0; grp < ngroups; ++grp) {
            const int gEnd = c_groupStart[grp + 1];
            do_group(g, gEnd);
            g = gEnd;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// regular (not-neighbour) pairs, the dominant kernel: grouped evaluation (see i2_pair.cuh, point_terms)
//   * one thread = one task at level 0 (G = min(4^level, 32) lanes per task otherwise);
//   * the 13 Gauss points of the (child) control panel are staged in shared memory, [point][component][thread],
//     so the loop keeps only the influence triangle (21 doubles) and the running products in registers;
//   * per point: lengths, 6 log arguments, (num, den) of the solid angle -> running products;
//     per group of equal weights: 3 x log_ratio + 1 x atan2_fast;
//   * control flow is warp-uniform: the only data-dependent decision (angle-sum overflow of a group) is taken by
//     __all_sync over the warp.
// ---------------------------------------------------------------------------------------------------------
// VAR bit 0: EDGELEN variant of point_terms; bit 1: primitives without the residual correction; bit 2: LEVEL0
// specialisation (no child loop, no per-level area scaling kept live)
template <int MINB, int VAR>
__global__ void __launch_bounds__(kThreads, MINB)
k_regular_grouped(PackedMesh pm, const int *__restrict__ tasks, const int *__restrict__ list, const int *__restrict__ countDev,
                  long long countHost, long long half, int level, int flags, double *__restrict__ out, double *__restrict__ results) {
    __shared__ double smM[MAX_GAUSS_POINTS * 3 * kThreads];
    __shared__ double smJ[kThreads / 32][96];   // one warp's 32 results, staged for 16-byte coalesced stores
    const long long count = countDev ? (long long)*countDev : countHost;
    constexpr bool EDGELEN = (VAR & 1) != 0, RESID = (VAR & 2) == 0, LEVEL0 = (VAR & 4) != 0, DERIVE = (VAR & 8) != 0 && EDGELEN;
    constexpr bool PROJ = (VAR & 16) != 0 && DERIVE;
    constexpr int UNR = (VAR & 32) ? 1 : kPointUnroll;
    constexpr int FIX = (VAR & 64) ? 13 : 0;
    constexpr bool LEVEL1 = (VAR & 128) != 0 && !LEVEL0;   // round 1 of the work queue: 4 lanes per task, one child each
    if (LEVEL0) level = 0;
    if (LEVEL1) level = 1;
    const LaneLayout lay(level);
    const int G = lay.G, perLane = lay.perLane;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ng = c_ngauss;
    const int stride = pm.stride;
    const double *__restrict__ tri = pm.tri;
    double *myM = smM + threadIdx.x;
    __shared__ double red[kThreads / 32][4];

    // per-lane partial (Psi, Theta) of task (i, j) over this lane's children; iStaged remembers whose Gauss points sit in smem
    // list-driven rounds with two tasks per warp: each task's 16 lanes vote on their own (see grouped_eval)
    const unsigned vmask = (list && G == 16) ? (lane < 16 ? 0x0000ffffu : 0xffff0000u) : 0xffffffffu;
    auto lane_work = [&](int i, int j, int sub, int &iStaged) -> d4 {
        double Si = __ldg(tri + PK_S * stride + i);
        for (int l = 0; l < level; ++l) Si *= 0.25;
        TriJ T;
        T.A = ld3(tri + PK_A * stride, stride, j);
        if (!DERIVE) { T.B = ld3(tri + PK_B * stride, stride, j); T.C = ld3(tri + PK_C * stride, stride, j); }
        T.ta = ld3(tri + PK_TA * stride, stride, j); T.tb = ld3(tri + PK_TB * stride, stride, j); T.tc = ld3(tri + PK_TC * stride, stride, j);
        T.Nu = ld3(tri + PK_NU * stride, stride, j);
        if (EDGELEN) { const d3 L = ld3(tri + PK_L * stride, stride, j); T.La = L.x; T.Lb = L.y; T.Lc = L.z; }
        if (PROJ) {
            T.c2 = T.Lc * T.Lb * dot(T.tc, T.tb);
            T.s1 = hi_word(T.Lc) - kNearEdgeShift; T.s2 = hi_word(T.La) - kNearEdgeShift; T.s3 = hi_word(T.Lb) - kNearEdgeShift;
        }
        double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
        for (int k = 0; k < perLane; ++k) {
            if (perLane > 1 || i != iStaged) {
                d3 A = ld3(tri + PK_A * stride, stride, i), B = ld3(tri + PK_B * stride, stride, i), C = ld3(tri + PK_C * stride, stride, i);
                descend(A, B, C, level, sub + G * k);
#pragma unroll 1
                for (int g = 0; g < ng; ++g) {
                    const d3 M = gauss_point(g, A, B, C);
                    myM[(3 * g + 0) * kThreads] = M.x; myM[(3 * g + 1) * kThreads] = M.y; myM[(3 * g + 2) * kThreads] = M.z;
                }
                iStaged = perLane > 1 ? -1 : i;
            }
            double a1, a2, a3, a4;
            grouped_eval<EDGELEN, RESID, DERIVE, PROJ, kThreads, UNR, FIX>(myM, ng, T, a1, a2, a3, a4, vmask);
            if (LEVEL0 || LEVEL1) { s1 = Si * a1; s2 = Si * a2; s3 = Si * a3; s4 = Si * a4; }
            else { s1 = fma(Si, a1, s1); s2 = fma(Si, a2, s2); s3 = fma(Si, a3, s3); s4 = fma(Si, a4, s4); }
        }
        return vec4(s1 * T.tc + s2 * T.ta + s3 * T.tb, s4);
    };
    auto store = [&](int slot, int j, d4 total) {
        double2 *o = reinterpret_cast<double2 *>(out + 4 * (long long)slot);
        o[0] = make_double2(total.x, total.y);
        o[1] = make_double2(total.z, total.w);
        if (results) {   // fixed level: final assembly fused (no second pass over the integrals)
            const d3 J = assemble_J(total, ld3(tri + PK_N * stride, stride, j), 0.0, false);
            results[3 * (long long)slot] = J.x; results[3 * (long long)slot + 1] = J.y; results[3 * (long long)slot + 2] = J.z;
        }
    };
    // the same for a warp whose 32 lanes hold the 32 consecutive slots slot0 .. slot0+31 (level 0, no list): the 768 bytes of
    // results leave as 48 16-byte stores instead of 96 strided 8-byte ones — full sectors, which is what a store into another
    // GPU's memory over NVLink (multi-GPU export, i2_peer_*) needs to reach the link's bandwidth
    auto store_warp = [&](int slot0, int j, d4 total) {
        double2 *o = reinterpret_cast<double2 *>(out + 4 * (long long)(slot0 + lane));
        o[0] = make_double2(total.x, total.y);
        o[1] = make_double2(total.z, total.w);
        const d3 J = assemble_J(total, ld3(tri + PK_N * stride, stride, j), 0.0, false);
        double *sj = smJ[warp];
        sj[3 * lane] = J.x; sj[3 * lane + 1] = J.y; sj[3 * lane + 2] = J.z;
        __syncwarp();
        double *dst = results + 3 * (long long)slot0;
        const int mis = (int)((reinterpret_cast<unsigned long long>(dst) >> 3) & 1ull);   // first element not 16-byte aligned: peel it
#pragma unroll
        for (int k = lane; k < 48; k += 32) {
            const int e = mis + 2 * k;
            if (e + 1 < 96) *reinterpret_cast<double2 *>(dst + e) = make_double2(sj[e], sj[e + 1]);
        }
        if (mis && lane == 0) dst[0] = sj[0];
        if (mis && lane == 1) dst[95] = sj[95];
        __syncwarp();
    };
    const bool vecStores = results && !list && G == 1 && (flags & 1);

    if (G <= 32) {
        // A warp walks a contiguous chunk of kChunkIters x (32/G) tasks: lists are sorted by control panel i, so a lane meets
        // the same i (and the same child) in consecutive iterations and its staged Gauss points are reused (13 x 9 FP64 per
        // task saved); chunks are dealt round-robin to the warps of the grid.
        // (list-driven rounds of the work queue have few tasks and no sorted control panels to reuse: one iteration per chunk, so
        // that the tasks spread over all warps of the grid instead of queueing 16 deep behind a few of them)
        const int kChunkIters = list ? 1 : 16;
        const int groupsPerWarp = 32 / G, sub = lane & (G - 1);
        // 32-bit index arithmetic throughout (task counts fit an int: the reference's int3 slots); unsigned for the padded positions
        const unsigned warpId = (blockIdx.x * kThreads + threadIdx.x) >> 5, warpStride = (gridDim.x * kThreads) >> 5;
        const unsigned chunkTasks = (unsigned)groupsPerWarp * kChunkIters;
        // `half` > 0: the list is two segments, slots [0, half) and [half, count) (pairs and reversed pairs of
        // Evaluator3D::runAllPairs); the second segment starts a new warp iteration, so which tasks decide together depends
        // only on a task's position inside its segment — a shard made of 32-aligned pieces of both segments
        // (i2_host_set_shard, i2_mgpu_*) reproduces the unsharded run bit for bit.
        const int cnt = (int)count, H = list ? 0 : (int)half;
        const unsigned P1 = (unsigned)(H + groupsPerWarp - 1) / groupsPerWarp * groupsPerWarp;   // first segment, padded to whole iterations
        const unsigned padded = P1 + (unsigned)(cnt - H);
        const int laneTask = lane / G;
        int iStaged = -1;
        for (unsigned chunk = warpId; (unsigned long long)chunk * chunkTasks < padded; chunk += warpStride) {
            const unsigned chunkBase = chunk * chunkTasks;
            for (int it = 0; it < kChunkIters; ++it) {
                const unsigned base = chunkBase + (unsigned)(it * groupsPerWarp);
                if (base >= padded) break;
                // a warp iteration never straddles the two segments (P1 is a multiple of the tasks per iteration)
                const bool first = base < P1;
                const int segEnd = first ? H : cnt;
                int r = (int)(first ? base : base - P1 + (unsigned)H) + laneTask;
                const bool active = r < segEnd;
                if (!active) r = segEnd - 1;   // tail lanes recompute the segment's last task (no write) so that warp votes stay full-mask
                const int slot = list ? __ldg(list + r) : r;
                const int i = __ldg(tasks + 3 * (long long)slot), j = __ldg(tasks + 3 * (long long)slot + 1);
                const d4 total = warp_sum(lane_work(i, j, sub, iStaged), G);
                if (vecStores && __all_sync(0xffffffffu, active)) store_warp(slot - lane, j, total);
                else if (active && sub == 0) store(slot, j, total);
            }
        }
    } else {
        // deep refinement: the lanes of one task span G/32 warps of the CTA (see LaneLayout)
        const int tasksPerCTA = kThreads / G, sub = threadIdx.x & (G - 1), tIdx = threadIdx.x / G, warpsPerTask = G / 32;
        for (long long base = (long long)blockIdx.x * tasksPerCTA; base < count; base += (long long)gridDim.x * tasksPerCTA) {
            long long r = base + tIdx;
            const bool active = r < count;
            if (!active) r = count - 1;
            const int slot = list ? __ldg(list + r) : (int)r;
            const int i = __ldg(tasks + 3 * (long long)slot), j = __ldg(tasks + 3 * (long long)slot + 1);
            int iStaged = -1;
            const d4 part = warp_sum(lane_work(i, j, sub, iStaged), 32);
            if (lane == 0) { red[warp][0] = part.x; red[warp][1] = part.y; red[warp][2] = part.z; red[warp][3] = part.w; }
            __syncthreads();
            if (active && sub == 0) {
                d4 total = {0.0, 0.0, 0.0, 0.0};
                for (int w = 0; w < warpsPerTask; ++w) { total.x += red[warp + w][0]; total.y += red[warp + w][1]; total.z += red[warp + w][2]; total.w += red[warp + w][3]; }
                store(slot, j, total);
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// list-free ("matrix-free") regular class: out[i] = sum over all j that share no vertex with i of w_j J(K_i, K_j).
// No task list and no per-pair result is materialised, so meshes beyond the reference's int32 / N^2-list limits
// (N > 46 340 triangles) fit: the classification IS the vertex-id comparison done here on the fly.
//   thread = one control panel i (its Gauss points staged in shared memory, its row sum in registers);
//   CTA    = 128 rows x one chunk of columns; columns stream through shared memory in tiles of kTileJ
//            (SoA -> smem coalesced, then every thread reads the same j: broadcast);
//   grid.y = column chunks; per-chunk partial row sums are written out and reduced in fixed order (deterministic).
// Pairs that share a vertex (and i == j) are evaluated like the others and discarded by a select, so the warp votes of
// grouped_eval stay full-mask and the j loop is uniform.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTileJ = 32;
constexpr int kTileStride = 29;   // 24 geometry + 3 normal + 1 weight + 1 pad(ids live in a separate int array)

// TILE = columns per shared-memory tile (32, or 16 so that five CTAs fit the SM's shared memory: MINB = 5)
template <int MINB, int TILE = kTileJ>
__global__ void __launch_bounds__(kThreads, MINB)
k_apply_regular(PackedMesh pm, int rowLo, int rowHi, int colLo, int colHi, int colChunk, const double *__restrict__ weights,
                double *__restrict__ partial) {
    constexpr int kTileJ = TILE;   // shadows the namespace constant inside this kernel
    __shared__ double smM[MAX_GAUSS_POINTS * 3 * kThreads];
    __shared__ double smT[kTileJ * kTileStride];
    __shared__ int smId[kTileJ * 3];
    const int rows = rowHi - rowLo;
    const int r = blockIdx.x * kThreads + threadIdx.x;
    const bool active = r < rows;
    const int i = rowLo + (active ? r : rows - 1);
    const int ng = c_ngauss;
    const int stride = pm.stride;
    const double *__restrict__ tri = pm.tri;
    double *myM = smM + threadIdx.x;
    const tri3 ci = ldtri(pm.cells, i);
    const double Si = __ldg(tri + PK_S * stride + i);
    {
        const d3 A = ld3(tri + PK_A * stride, stride, i), B = ld3(tri + PK_B * stride, stride, i), C = ld3(tri + PK_C * stride, stride, i);
#pragma unroll 1
        for (int g = 0; g < ng; ++g) {
            const d3 M = gauss_point(g, A, B, C);
            myM[(3 * g + 0) * kThreads] = M.x; myM[(3 * g + 1) * kThreads] = M.y; myM[(3 * g + 2) * kThreads] = M.z;
        }
    }
    d3 acc = {0.0, 0.0, 0.0};
    const int j0 = colLo + blockIdx.y * colChunk;
    const int j1 = min(j0 + colChunk, colHi);
    for (int tile = j0; tile < j1; tile += kTileJ) {
        __syncthreads();
        // cooperative, coalesced load of the tile: component-major source rows -> [j][component] in shared memory
        for (int e = threadIdx.x; e < kTileJ * 28; e += kThreads) {
            const int comp = e / kTileJ, jj = e % kTileJ, j = tile + jj;
            double v = 0.0;
            if (j < j1) {
                if (comp < 21) v = __ldg(tri + comp * stride + j);                       // A,B,C,ta,tb,tc,Nu
                else if (comp < 24) v = __ldg(tri + (PK_L + comp - 21) * stride + j);     // edge lengths
                else if (comp < 27) v = __ldg(tri + (PK_N + comp - 24) * stride + j);     // unit normal
                else v = weights ? __ldg(weights + j) : 1.0;
            }
            smT[jj * kTileStride + comp] = v;
        }
        for (int e = threadIdx.x; e < kTileJ * 3; e += kThreads) {
            const int jj = e / 3, j = tile + jj;
            smId[e] = j < j1 ? __ldg(pm.cells + 3 * (long long)j + e % 3) : -1;
        }
        __syncthreads();
        const int nj = min(kTileJ, j1 - tile);
#pragma unroll 1
        for (int jj = 0; jj < nj; ++jj) {
            const double *t = smT + jj * kTileStride;
            TriJ T;
            T.A = {t[0], t[1], t[2]};   // B and C are derived from A, the tangents and the edge lengths (DERIVE)
            T.ta = {t[9], t[10], t[11]}; T.tb = {t[12], t[13], t[14]}; T.tc = {t[15], t[16], t[17]};
            T.Nu = {t[18], t[19], t[20]};
            T.La = t[21]; T.Lb = t[22]; T.Lc = t[23];
            T.c2 = T.Lc * T.Lb * dot(T.tc, T.tb);
            T.s1 = hi_word(T.Lc) - kNearEdgeShift; T.s2 = hi_word(T.La) - kNearEdgeShift; T.s3 = hi_word(T.Lb) - kNearEdgeShift;
            double a1, a2, a3, a4;
            grouped_eval<true, false, true, true>(myM, ng, T, a1, a2, a3, a4);
            const int ja = smId[3 * jj], jb = smId[3 * jj + 1], jc = smId[3 * jj + 2];
            const bool skip = (tile + jj == i) || ci.a == ja || ci.a == jb || ci.a == jc || ci.b == ja || ci.b == jb || ci.b == jc ||
                              ci.c == ja || ci.c == jb || ci.c == jc;
            const d3 psi = (Si * a1) * T.tc + (Si * a2) * T.ta + (Si * a3) * T.tb;
            const d3 nj3 = {t[24], t[25], t[26]};
            const d3 J = assemble_J(vec4(psi, Si * a4), nj3, 0.0, false);
            const double w = t[27];
            acc.x = skip ? acc.x : fma(w, J.x, acc.x);
            acc.y = skip ? acc.y : fma(w, J.y, acc.y);
            acc.z = skip ? acc.z : fma(w, J.z, acc.z);
        }
    }
    if (active) {
        double *o = partial + ((long long)blockIdx.y * rows + r) * 3;
        o[0] = acc.x; o[1] = acc.y; o[2] = acc.z;
    }
}

__global__ void k_reduce_partials(const double *__restrict__ partial, int rows, int chunks, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * 3) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += partial[(long long)c * rows * 3 + e];
    out[e] = s;
}

void launch_apply_regular(const PackedMesh &pm, int rowLo, int rowHi, int colLo, int colHi, int chunks, const double *weights,
                          double *partial, double *out3, cudaStream_t s) {
    const int rows = rowHi - rowLo, cols = colHi - colLo;
    if (rows <= 0 || cols <= 0) return;
    const int colChunk = (cols + chunks - 1) / chunks;
    dim3 grid((rows + kThreads - 1) / kThreads, chunks);
    ++g_launchCount;
    // 94 registers without spills and 16-column tiles (44 KB of shared memory) let five CTAs share an SM: 849 ms against
    // 886 ms with four (126 registers, 32-column tiles) on the 108 544-triangle sphere, same bits
    // (profiles/r02_ab_apply_minb.log).  env I2_APPLY_MINB=4 restores the four-CTA variant.
    static const int minb = [] { const char *e = getenv("I2_APPLY_MINB"); return e ? atoi(e) : 5; }();
    if (minb == 5) k_apply_regular<5, 16><<<grid, kThreads, 0, s>>>(pm, rowLo, rowHi, colLo, colHi, colChunk, weights, partial);
    else k_apply_regular<4><<<grid, kThreads, 0, s>>>(pm, rowLo, rowHi, colLo, colHi, colChunk, weights, partial);
    ++g_launchCount;
    k_reduce_partials<<<(rows * 3 + 255) / 256, 256, 0, s>>>(partial, rows, chunks, out3);
}

// ---------------------------------------------------------------------------------------------------------
// list-free regular class WITH automatic error control: the Runge loop of NumericalIntegrator3D / EvaluatorJ3DK
// (src/evaluators/evaluatorJ3DK.cu:895-1012, src/evaluators/evaluator3d.cu:76-99) per pair, without a task list, a
// refined mesh or a work queue in memory — for meshes beyond the reference's N^2-list limits.
//   4 lanes = one control panel i; lane l holds the Gauss points of child l of i (level 1) in shared memory, the
//   parent's points sit once per row.  Columns stream through shared memory in tiles; per step of 4 columns every lane
//   evaluates its child against the 4 triangles (-> I_1 by a 4-lane shuffle sum, round 1) and the parent against ONE of
//   them (-> I_0, round 0): 5 evaluations per lane per 4 pairs, perfectly balanced.  Lane q then owns pair (i, j_q):
//   Runge criterion; a pair that fails it (rare: < 1 % of the regular pairs of the example meshes) is refined further by
//   the WHOLE warp, 32 lanes over the 16 / 64 / 256 / 1024 children of rounds 2..5, until it converges.
//   Everything is summed in a fixed order: results are bitwise reproducible.
// The reference keeps converged values in two ping-pong buffers without copying: the final value of a
// pair is the newest value of the buffer selected by the parity of the class's LAST round L.  Both candidates are
// accumulated (accE: newest even-round value, accO: newest odd-round value); k_reduce_partials_adaptive picks by L.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAdRows = kThreads / 4;
constexpr int kAdTileJ = 16;
constexpr size_t kAdaptiveSmemBytes = sizeof(double) * (MAX_GAUSS_POINTS * 3 * (kThreads + kAdRows) + kAdTileJ * kTileStride);

struct ApplyAdaptiveOut {
    double *partial;             // [chunks][rows][6]: accE.xyz, accO.xyz
    unsigned char *depth;        // [chunks][rows]: max over the chunk's pairs of the number of compare rounds failed (0..5)
    int *lastRound;              // L of the class (atomicMax)
    unsigned long long *counts;  // [6]: counts[0] = pairs examined, counts[m] = pairs unconverged after round m
};

static __device__ __forceinline__ d4 shfl4(d4 v, int src) {
    return {__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src),
            __shfl_sync(0xffffffffu, v.w, src)};
}

// influence triangle from one row of a shared-memory column tile (B and C are derived from A, the tangents and the edge lengths)
static __device__ __forceinline__ void load_tile_T(const double *t, TriJ &T) {
    T.A = {t[0], t[1], t[2]};
    T.ta = {t[9], t[10], t[11]}; T.tb = {t[12], t[13], t[14]}; T.tc = {t[15], t[16], t[17]};
    T.Nu = {t[18], t[19], t[20]};
    T.La = t[21]; T.Lb = t[22]; T.Lc = t[23];
    T.c2 = T.Lc * T.Lb * dot(T.tc, T.tb);
    T.s1 = hi_word(T.Lc) - kNearEdgeShift; T.s2 = hi_word(T.La) - kNearEdgeShift; T.s3 = hi_word(T.Lb) - kNearEdgeShift;
}
// Gauss points of child `c` (level `lev`) of control panel `ii` into the calling thread's slot myC ([point][component], stride kThreads)
static __device__ __forceinline__ void stage_child_points(const double *__restrict__ tri, int stride, int ng, int ii, int lev, int c, double *myC) {
    d3 A = ld3(tri + PK_A * stride, stride, ii), B = ld3(tri + PK_B * stride, stride, ii), C = ld3(tri + PK_C * stride, stride, ii);
    descend(A, B, C, lev, c);
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
        const d3 M = gauss_point(g, A, B, C);
        myC[(3 * g + 0) * kThreads] = M.x; myC[(3 * g + 1) * kThreads] = M.y; myC[(3 * g + 2) * kThreads] = M.z;
    }
}
// Rounds 2..5 of ONE pair (control panel srcI of area srcS against the tile row tT), all 32 lanes of the warp over the
// 16 / 64 / 256 / 1024 children.  I1 / I0 = the pair's round-1 / round-0 values.  Rare path, kept out of line so that its
// registers do not weigh on the main loop.  res: newest even-round value, newest odd-round value; last = last round
// executed, failed = number of compare rounds the pair failed.
struct DeepResult { d4 evenV, oddV; int last, failed; };
static __device__ __noinline__ void deep_rounds(const double *__restrict__ tri, int stride, int ng, double *myC, const double *tT, int srcI,
                                                double srcS, d4 I1, d4 I0, double pow2p, unsigned int *smCount, DeepResult *res) {
    const int lane = threadIdx.x & 31;
    TriJ Ts;
    load_tile_T(tT, Ts);
    d4 prev = I1, evenV = I0, oddV = I1;
    int m = 2;
    bool un = true;
    double Sm = 0.0625 * srcS;
#pragma unroll 1
    for (; m <= MAX_REFINE_LEVEL && un; ++m, Sm *= 0.25) {
        const int children = 1 << (2 * m);
        const int perLane = children > 32 ? children / 32 : 1;
        d4 part = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int k = 0; k < perLane; ++k) {
            const int c = lane + 32 * k;
            stage_child_points(tri, stride, ng, srcI, m, c & (children - 1), myC);   // lanes beyond the 16 children of round 2 duplicate one
            const double S = c < children ? Sm : 0.0;                                 // ... with weight 0 (the votes stay full-mask)
            double a1, a2, a3, a4;
            grouped_eval<true, false, true, true>(myC, ng, Ts, a1, a2, a3, a4);
            const d4 e = vec4((S * a1) * Ts.tc + (S * a2) * Ts.ta + (S * a3) * Ts.tb, S * a4);
            part.x += e.x; part.y += e.y; part.z += e.z; part.w += e.w;
        }
        const d4 cur = warp_sum(part, 32);
        un = runge_unconverged(cur, prev, pow2p);
        if (m & 1) oddV = cur; else evenV = cur;
        prev = cur;
        if (un && lane == 0) atomicAdd(&smCount[m], 1u);
    }
    res->evenV = evenV; res->oddV = oddV; res->last = m - 1; res->failed = un ? MAX_REFINE_LEVEL : m - 2;
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_apply_regular_adaptive(PackedMesh pm, int rowLo, int rowHi, int colLo, int colHi, int colChunk, const double *__restrict__ weights,
                         ApplyAdaptiveOut o) {
    extern __shared__ double smDyn[];                                   // 53.6 KB: above the 48 KB static limit
    double *smC = smDyn;                                                  // per thread: Gauss points of its child (or of a deep child)
    double *smP = smC + MAX_GAUSS_POINTS * 3 * kThreads;                  // per row: Gauss points of the control panel itself
    double *smT = smP + MAX_GAUSS_POINTS * 3 * kAdRows;                   // column tile
    __shared__ int smId[kAdTileJ * 3];
    __shared__ unsigned int smCount[6];
    const int rows = rowHi - rowLo;
    const int lane = threadIdx.x & 31, sub = threadIdx.x & 3, rowInCta = threadIdx.x >> 2;
    const int r = blockIdx.x * kAdRows + rowInCta;
    const bool active = r < rows;
    const int i = rowLo + (active ? r : rows - 1);
    const int ng = c_ngauss;
    const int stride = pm.stride;
    const double pow2p = c_pow2p;
    const double *__restrict__ tri = pm.tri;
    double *myC = smC + threadIdx.x;
    const double *myP = smP + rowInCta;
    const tri3 ci = ldtri(pm.cells, i);
    const double Si = __ldg(tri + PK_S * stride + i);
    if (threadIdx.x < 6) smCount[threadIdx.x] = 0;

    auto stage_child = [&](int ii, int lev, int c) { stage_child_points(tri, stride, ng, ii, lev, c, myC); };
    auto load_T = [&](int jj, TriJ &T) { load_tile_T(smT + jj * kTileStride, T); };
    // (Psi, Theta) of one staged panel of area S against T
    auto eval_child = [&](const TriJ &T, double S) -> d4 {
        double a1, a2, a3, a4;
        grouped_eval<true, false, true, true>(myC, ng, T, a1, a2, a3, a4);
        return vec4((S * a1) * T.tc + (S * a2) * T.ta + (S * a3) * T.tb, S * a4);
    };

    {   // the four lanes of a row share the work of staging the parent's points; each stages its own child
        const d3 A = ld3(tri + PK_A * stride, stride, i), B = ld3(tri + PK_B * stride, stride, i), C = ld3(tri + PK_C * stride, stride, i);
        for (int g = sub; g < ng; g += 4) {
            const d3 M = gauss_point(g, A, B, C);
            smP[(3 * g + 0) * kAdRows + rowInCta] = M.x; smP[(3 * g + 1) * kAdRows + rowInCta] = M.y; smP[(3 * g + 2) * kAdRows + rowInCta] = M.z;
        }
        stage_child(i, 1, sub);
    }

    d3 accE = {0.0, 0.0, 0.0}, accO = {0.0, 0.0, 0.0};
    int depth = 0, lastRound = 0;
    unsigned int examined = 0;
    const int j0 = colLo + blockIdx.y * colChunk;
    const int j1 = min(j0 + colChunk, colHi);
    for (int tile = j0; tile < j1; tile += kAdTileJ) {
        __syncthreads();
        for (int e = threadIdx.x; e < kAdTileJ * 28; e += kThreads) {
            const int comp = e / kAdTileJ, jj = e % kAdTileJ, j = tile + jj;
            double v = 0.0;
            if (j < j1) {
                if (comp < 21) v = __ldg(tri + comp * stride + j);
                else if (comp < 24) v = __ldg(tri + (PK_L + comp - 21) * stride + j);
                else if (comp < 27) v = __ldg(tri + (PK_N + comp - 24) * stride + j);
                else v = weights ? __ldg(weights + j) : 1.0;
            }
            smT[jj * kTileStride + comp] = v;
        }
        for (int e = threadIdx.x; e < kAdTileJ * 3; e += kThreads) {
            const int jj = e / 3, j = tile + jj;
            smId[e] = j < j1 ? __ldg(pm.cells + 3 * (long long)j + e % 3) : -1;
        }
        __syncthreads();
        const int nj = min(kAdTileJ, j1 - tile);
#pragma unroll 1
        for (int jj0 = 0; jj0 < nj; jj0 += 4) {
            // round 1: every lane's child against the 4 triangles of the step; lane q keeps the 4-lane sum for column q
            d4 I1 = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                TriJ T;
                load_T(min(jj0 + q, nj - 1), T);   // a clamped duplicate is discarded with its owner lane below
                const d4 c = warp_sum(eval_child(T, 0.25 * Si), 4);
                if (q == sub) I1 = c;
            }
            // round 0: the control panel itself against this lane's column
            const int jm = min(jj0 + sub, nj - 1);
            TriJ T;
            load_T(jm, T);
            d4 I0;
            {
                double a1, a2, a3, a4;
                grouped_eval<true, false, true, true, kAdRows>(myP, ng, T, a1, a2, a3, a4);
                I0 = vec4((Si * a1) * T.tc + (Si * a2) * T.ta + (Si * a3) * T.tb, Si * a4);
            }
            const int ja = smId[3 * jm], jb = smId[3 * jm + 1], jc = smId[3 * jm + 2];
            const bool skip = !active || (jj0 + sub >= nj) || (tile + jm == i) || ci.a == ja || ci.a == jb || ci.a == jc ||
                              ci.b == ja || ci.b == jb || ci.b == jc || ci.c == ja || ci.c == jb || ci.c == jc;
            bool unconv = !skip && runge_unconverged(I1, I0, pow2p);
            d4 valE = I0, valO = I1;      // newest even-round / odd-round value of this lane's pair
            int myLast = 1, myFailed = unconv ? 1 : 0;
            unsigned um = __ballot_sync(0xffffffffu, unconv);
            if (um) {
                if (lane == 0) atomicAdd(&smCount[1], __popc(um));
                // rounds 2..5 of every failing pair, one pair at a time, 32 lanes over its children (warp-uniform loop)
                while (um) {
                    const int src = __ffs(um) - 1;
                    um &= um - 1;
                    const int srcI = __shfl_sync(0xffffffffu, i, src), srcJm = __shfl_sync(0xffffffffu, jm, src);
                    const double srcS = __shfl_sync(0xffffffffu, Si, src);
                    DeepResult res;
                    deep_rounds(tri, stride, ng, myC, smT + srcJm * kTileStride, srcI, srcS, shfl4(I1, src), shfl4(I0, src), pow2p, smCount, &res);
                    if (lane == src) { valE = res.evenV; valO = res.oddV; myLast = res.last; myFailed = res.failed; }
                }
                stage_child(i, 1, sub);   // the deep rounds overwrote this thread's slot
            }
            if (!skip) {
                const d3 nj3 = {smT[jm * kTileStride + 24], smT[jm * kTileStride + 25], smT[jm * kTileStride + 26]};
                const double w = smT[jm * kTileStride + 27];
                const d3 JE = assemble_J(valE, nj3, 0.0, false), JO = assemble_J(valO, nj3, 0.0, false);
                accE.x = fma(w, JE.x, accE.x); accE.y = fma(w, JE.y, accE.y); accE.z = fma(w, JE.z, accE.z);
                accO.x = fma(w, JO.x, accO.x); accO.y = fma(w, JO.y, accO.y); accO.z = fma(w, JO.z, accO.z);
                depth = max(depth, myFailed);
                lastRound = max(lastRound, myLast);
                ++examined;
            }
        }
    }
    // the 4 lanes of a row hold the sums of their columns: fixed-order 4-lane sum, lane 0 of the row writes
    d4 e4 = warp_sum(vec4(accE, 0.0), 4), o4 = warp_sum(vec4(accO, 0.0), 4);
    for (int off = 2; off > 0; off >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, off));
    if (active && sub == 0) {
        double *p = o.partial + ((long long)blockIdx.y * rows + r) * 6;
        p[0] = e4.x; p[1] = e4.y; p[2] = e4.z; p[3] = o4.x; p[4] = o4.y; p[5] = o4.z;
        o.depth[(long long)blockIdx.y * rows + r] = (unsigned char)depth;
    }
    lastRound = __reduce_max_sync(0xffffffffu, lastRound);
    examined = __reduce_add_sync(0xffffffffu, examined);
    if (lane == 0) {
        if (lastRound > 0) atomicMax(o.lastRound, lastRound);
        atomicAdd(o.counts, (unsigned long long)examined);
    }
    __syncthreads();
    if (threadIdx.x >= 1 && threadIdx.x < 6 && smCount[threadIdx.x]) atomicAdd(o.counts + threadIdx.x, (unsigned long long)smCount[threadIdx.x]);
}

// out[r] = sum over chunks of the candidate selected by the parity of the class's last round L (odd L: newest odd-round
// values), other[r] = the other candidate (a multi-GPU caller whose global L has the other parity takes that one);
// refinements[r] = 1 + max over the row's pairs of the compare rounds failed (NumericalIntegrator3D::getRefinementsRequired)
__global__ void k_reduce_partials_adaptive(const double *__restrict__ partial, const unsigned char *__restrict__ depth, int rows, int chunks,
                                           const int *__restrict__ lastRound, double *__restrict__ out, double *__restrict__ other,
                                           unsigned char *__restrict__ refinements) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * 3) return;
    const int r = e / 3, k = e % 3;
    const bool odd = (*lastRound & 1) != 0;
    double sE = 0.0, sO = 0.0;
    for (int c = 0; c < chunks; ++c) {
        const double *p = partial + ((long long)c * rows + r) * 6;
        sE += p[k]; sO += p[3 + k];
    }
    out[e] = odd ? sO : sE;
    if (other) other[e] = odd ? sE : sO;
    if (refinements && k == 0) {
        int d = 0;
        for (int c = 0; c < chunks; ++c) d = max(d, (int)depth[(long long)c * rows + r]);
        refinements[r] = (unsigned char)(1 + d);
    }
}

void launch_apply_regular_adaptive(const PackedMesh &pm, int rowLo, int rowHi, int colLo, int colHi, int chunks, const double *weights,
                                   double *partial6, unsigned char *depth, int *lastRound, unsigned long long *counts6, double *out3,
                                   double *other3, unsigned char *refinements, cudaStream_t s) {
    const int rows = rowHi - rowLo, cols = colHi - colLo;
    if (rows <= 0 || cols <= 0) return;
    const int colChunk = (cols + chunks - 1) / chunks;
    dim3 grid((rows + kAdRows - 1) / kAdRows, chunks);
    ApplyAdaptiveOut o{partial6, depth, lastRound, counts6};
    constexpr size_t smem = kAdaptiveSmemBytes;
    ++g_launchCount;
    // four CTAs per SM at 128 registers: 441 ms against 451 ms with three (168 registers) on s5m2 refined once, same bits
    // (profiles/r02_ab_apply_minb.log) — the extra spill traffic sits in the rare deep rounds.  env I2_APPLY_AD_MINB=3 restores.
    static const int minb = [] { const char *e = getenv("I2_APPLY_AD_MINB"); return e ? atoi(e) : 4; }();
    if (minb == 4) k_apply_regular_adaptive<4><<<grid, kThreads, smem, s>>>(pm, rowLo, rowHi, colLo, colHi, colChunk, weights, o);
    else k_apply_regular_adaptive<3><<<grid, kThreads, smem, s>>>(pm, rowLo, rowHi, colLo, colHi, colChunk, weights, o);
    if (out3) launch_reduce_partials_adaptive(partial6, depth, rows, chunks, lastRound, out3, other3, refinements, s);
}

// second half of the list-free adaptive pass: picks, per row, the candidate selected by the parity of *lastRound — a multi-GPU
// caller runs the first half with out3 == nullptr, agrees on the last round (maximum over the GPUs) and then calls this
void launch_reduce_partials_adaptive(const double *partial6, const unsigned char *depth, int rows, int chunks, const int *lastRound, double *out3,
                                     double *other3, unsigned char *refinements, cudaStream_t s) {
    if (rows <= 0) return;
    ++g_launchCount;
    k_reduce_partials_adaptive<<<(rows * 3 + 255) / 256, 256, 0, s>>>(partial6, depth, rows, chunks, lastRound, out3, other3, refinements);
}

// tuning knob (env I2_MINBLOCKS = 3|4|5): resident CTAs per SM the regular kernel is compiled for
static int g_minBlocks = [] { const char *e = getenv("I2_MINBLOCKS"); return e ? atoi(e) : 4; }();
// kernel flag bit 0: 16-byte coalesced result stores staged through shared memory.  Measured on Vint16k: 1.6 % SLOWER than the
// strided 8-byte stores when the results stay in local HBM (L2 merges the partial sectors anyway), so the host asks for them only
// when the results go to another GPU over NVLink (launch_integrate flags); env I2_VEC_STORES=1 forces them on (A/B knob)
static int g_kernelFlags = [] { const char *e = getenv("I2_VEC_STORES"); return (e && atoi(e) != 0) ? 1 : 0; }();
// the straight-line variant for rules with the group structure of Cowper's 13-point rule (launch flag bit 1, set by the context when
// upload_quadrature recognised the shape; env I2_FIX13=0 turns it off: A/B knob)
// MEASURED: 32.3 ms against 28.0 ms for the loop version on Vint16k (bit-identical results): the kernel grows from 24 KB to 80 KB of
// SASS and the four warps of a scheduler no longer share the instruction cache — smaller code wins.  Kept as an opt-in (I2_FIX13=1).
static int g_fix13 = [] { const char *e = getenv("I2_FIX13"); return (e && atoi(e) != 0) ? 1 : 0; }();
static int g_variant = [] { const char *e = getenv("I2_VARIANT"); return e ? atoi(e) : 27; }();

void launch_integrate(int cls, int mathMode, const PackedMesh &pm, const int *tasks, const int *list, const int *countDev,
                      long long countHost, long long half, int level, double *out4, double *fusedResults3, int numSMs, cudaStream_t s, int flags) {
    if (!countDev && countHost <= 0) return;
    const int children = 1 << (2 * level);
    const int G = children < kThreads ? children : kThreads;   // LaneLayout
    long long blocks;
    const long long persistent = (long long)numSMs * 4 * 8;  // a few waves of 4 CTAs/SM; the kernel grid-strides
    if (countDev) blocks = persistent;
    else {
        blocks = (countHost * G + kThreads - 1) / kThreads;
        if (blocks > persistent * 4) blocks = persistent * 4;
    }
    const unsigned gb = (unsigned)blocks;
    if (cls == 0) { ++g_launchCount; k_integrate<0, MATH_STRICT, 3><<<gb, kThreads, 0, s>>>(pm, tasks, list, countDev, countHost, level, out4); }
    else if (cls == 1) { ++g_launchCount; k_integrate<1, MATH_STRICT, 3><<<gb, kThreads, 0, s>>>(pm, tasks, list, countDev, countHost, level, out4); }
    else if (mathMode == MATH_STRICT) { ++g_launchCount; k_integrate<2, MATH_STRICT, 4><<<gb, kThreads, 0, s>>>(pm, tasks, list, countDev, countHost, level, out4); }
    else if (mathMode == MATH_FAST_LIBDEVICE) { ++g_launchCount; k_integrate<2, MATH_FAST_LIBDEVICE, 4><<<gb, kThreads, 0, s>>>(pm, tasks, list, countDev, countHost, level, out4); }
    else if (mathMode == MATH_FAST_POINTWISE) { ++g_launchCount; k_integrate<2, MATH_FAST_POINTWISE, 4><<<gb, kThreads, 0, s>>>(pm, tasks, list, countDev, countHost, level, out4); }
    else {
        // tuning knobs: I2_MINBLOCKS (3|4|5 resident CTAs/SM), I2_VARIANT (bit0 edge-length identity, bit1 no residual correction,
        // bit3 derive d_b, d_c from d_a instead of reading B and C, bit4 projection form of the lengths/dots, bit5 point loop not
        // unrolled; default 27 = bits 0,1,3,4);
        // the LEVEL0 specialisation (bit 2) is chosen automatically
        const int var = (g_variant & 3) | (level == 0 ? 4 : 0) | (g_variant & 56) | ((g_fix13 && (flags & 2)) ? 64 : 0);
        ++g_launchCount;
#define I2_LAUNCH_GROUPED(MB, V) k_regular_grouped<MB, V><<<gb, kThreads, 0, s>>>(pm, tasks, list, countDev, countHost, half, level, (flags & 1) | g_kernelFlags, out4, fusedResults3)
#define I2_PICK_VAR(MB)                                                                                              \
        switch (var) {                                                                                           \
        case 7: I2_LAUNCH_GROUPED(MB, 7); break; case 15: I2_LAUNCH_GROUPED(MB, 15); break;                     \
        case 27: I2_LAUNCH_GROUPED(MB, 27); break; case 31: I2_LAUNCH_GROUPED(MB, 31); break;                   \
        case 59: I2_LAUNCH_GROUPED(MB, 59); break; case 63: I2_LAUNCH_GROUPED(MB, 63); break;                   \
        default: if (level == 0) I2_LAUNCH_GROUPED(MB, 31); else I2_LAUNCH_GROUPED(MB, 27); break;                       \
        }
        static const int useLevel1 = [] { const char *e = getenv("I2_LEVEL1"); return (e && atoi(e) == 0) ? 0 : 1; }();   // A/B knob
        if (var & 64) { if (level == 0) I2_LAUNCH_GROUPED(4, 95); else I2_LAUNCH_GROUPED(4, 91); }
        else if (level == 1 && !list && useLevel1 && g_minBlocks == 4 && var == 27) I2_LAUNCH_GROUPED(4, 155);   // 27 | 128
        else if (g_minBlocks == 3) { I2_PICK_VAR(3) }
        else if (g_minBlocks == 5) { I2_PICK_VAR(5) }
        else { I2_PICK_VAR(4) }
#undef I2_PICK_VAR
#undef I2_LAUNCH_GROUPED
    }
}

// CUDA loads kernel images lazily on first launch (milliseconds for the large integrate kernels); touching them at context
// creation keeps that one-time cost out of the first timed integrateOver* call.
cudaError_t preload_rest();
cudaError_t preload_kernels() {
    cudaFuncAttributes a;
    cudaError_t e = preload_rest();
    if (e != cudaSuccess) return e;
    // Looking a kernel up is not enough: the first LAUNCH on a device still pays for loading the image (milliseconds for the large
    // integrate kernels, and the loads of several devices of one process serialise on a runtime lock: 8 GPUs driven by one
    // process spent 24 of the 28 ms of their first timed run there).  One empty launch of the hot kernels per context, here.
    {
        PackedMesh none{nullptr, nullptr, 0, 0};
        k_regular_grouped<4, 31><<<1, kThreads>>>(none, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr, nullptr);
        k_regular_grouped<4, 27><<<1, kThreads>>>(none, nullptr, nullptr, nullptr, 0, 0, 1, 0, nullptr, nullptr);
        k_integrate<0, MATH_STRICT, 3><<<1, kThreads>>>(none, nullptr, nullptr, nullptr, 0, 0, nullptr);
        k_integrate<1, MATH_STRICT, 3><<<1, kThreads>>>(none, nullptr, nullptr, nullptr, 0, 0, nullptr);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
#define I2_TOUCH(...) do { e = cudaFuncGetAttributes(&a, (const void *)(__VA_ARGS__)); if (e != cudaSuccess) return e; } while (0)
    I2_TOUCH(k_integrate<0, MATH_STRICT, 3>);
    I2_TOUCH(k_integrate<1, MATH_STRICT, 3>);
    I2_TOUCH(k_regular_grouped<4, 31>);
    I2_TOUCH(k_regular_grouped<4, 27>);
    I2_TOUCH(k_apply_regular<4>);
    I2_TOUCH((k_apply_regular<5, 16>));
    I2_TOUCH(k_apply_regular_adaptive<3>);
#undef I2_TOUCH
    e = cudaDeviceSynchronize();    // the empty launches above ran on the default stream
    if (e != cudaSuccess) return e;
    // function attributes are per device: opt in to > 48 KB of dynamic shared memory on the device of the calling context
    e = cudaFuncSetAttribute(k_apply_regular_adaptive<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAdaptiveSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_apply_regular_adaptive<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAdaptiveSmemBytes);
}

// ---------------------------------------------------------------------------------------------------------
// adaptive error control: Runge compare + warp-ballot compaction of the unconverged task slots
// ---------------------------------------------------------------------------------------------------------
// Runge comparison + DETERMINISTIC compaction of the unconverged slots.  Every CTA owns a contiguous segment of the input
// list and appends its unconverged slots, in input order, to the same segment of `staging` (ballot + prefix over the
// CTA's warps, no atomics); k_compact_gather then packs the segments densely in CTA order.  The output list therefore
// keeps the input order (ascending slots), whatever the scheduling: which tasks share a warp in the next round — and with
// it the warp-wide far-field tier decisions of the regular kernel — is the same in every run, so adaptive results are
// bitwise reproducible (an atomic ticket per warp made the order depend on which kernels ran concurrently).
constexpr int kCompareThreads = 256;
__global__ void __launch_bounds__(kCompareThreads)
k_compare(const double *__restrict__ cur, const double *__restrict__ prev, const int *__restrict__ tasks, const int *__restrict__ listIn,
          const int *__restrict__ countIn, long long countHost, int *__restrict__ staging, int *__restrict__ blockCnt, unsigned char *cellFlag,
          unsigned char *converged, QueueState *qs, int round) {
    __shared__ int warpCnt[kCompareThreads / 32];
    const long long n = countIn ? (long long)*countIn : countHost;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long seg = (n + gridDim.x - 1) / gridDim.x;
    const long long segLo = (long long)blockIdx.x * seg, segHi = segLo + seg < n ? segLo + seg : n;
    const double pow2p = c_pow2p;
    int running = 0;
    for (long long base = segLo; base < segHi; base += kCompareThreads) {
        const long long r = base + threadIdx.x;
        bool unconv = false;
        int slot = 0;
        if (r < segHi) {
            slot = listIn ? __ldg(listIn + r) : (int)r;
            const double2 *c = reinterpret_cast<const double2 *>(cur + 4 * (long long)slot);
            const double2 *p = reinterpret_cast<const double2 *>(prev + 4 * (long long)slot);
            const double2 c0 = c[0], c1 = c[1], p0 = p[0], p1 = p[1];
            unconv = runge_unconverged({c0.x, c0.y, c1.x, c1.y}, {p0.x, p0.y, p1.x, p1.y}, pow2p);
            if (converged) converged[slot] = unconv ? 0 : 1;
            if (unconv) cellFlag[__ldg(tasks + 3 * (long long)slot)] = 1;  // benign race: everyone writes 1
        }
        const unsigned m = __ballot_sync(0xffffffffu, unconv);
        if (lane == 0) warpCnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < kCompareThreads / 32; ++w) { const int cw = warpCnt[w]; before += w < warp ? cw : 0; total += cw; }
        if (unconv) staging[segLo + running + before + __popc(m & ((1u << lane) - 1u))] = slot;
        running += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) blockCnt[blockIdx.x] = running;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) qs->lastRound = round;
}

// packs the per-CTA segments of `staging` densely into listOut in CTA order; *countOut = total
__global__ void __launch_bounds__(kCompareThreads)
k_compact_gather(const int *__restrict__ staging, const int *__restrict__ blockCnt, const int *__restrict__ countIn, long long countHost,
                 int *__restrict__ listOut, int *countOut) {
    __shared__ int red[kCompareThreads / 32];
    __shared__ int offsetSh;
    const long long n = countIn ? (long long)*countIn : countHost;
    const long long seg = (n + gridDim.x - 1) / gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kCompareThreads) part += __ldg(blockCnt + b);
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int off = 0;
        for (int w = 0; w < kCompareThreads / 32; ++w) off += red[w];
        offsetSh = off;
    }
    __syncthreads();
    const int offset = offsetSh, mine = __ldg(blockCnt + blockIdx.x);
    const long long segLo = (long long)blockIdx.x * seg;
    for (int t = threadIdx.x; t < mine; t += kCompareThreads) listOut[offset + t] = staging[segLo + t];
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *countOut = offset + mine;
}

void launch_compare(const double *cur4, const double *prev4, const int *tasks, const int *listIn, const int *countIn, long long countHost,
                    int *staging, int *blockCnt, int *listOut, int *countOut, unsigned char *cellFlag, unsigned char *converged, QueueState *qs,
                    int round, int numSMs, cudaStream_t s) {
    long long blocks = countIn ? (long long)numSMs * 8 : (countHost + kCompareThreads - 1) / kCompareThreads;
    if (blocks > kCompareMaxBlocks) blocks = kCompareMaxBlocks;
    if (blocks < 1) blocks = 1;
    ++g_launchCount;
    k_compare<<<(unsigned)blocks, kCompareThreads, 0, s>>>(cur4, prev4, tasks, listIn, countIn, countHost, staging, blockCnt, cellFlag, converged, qs, round);
    ++g_launchCount;
    k_compact_gather<<<(unsigned)blocks, kCompareThreads, 0, s>>>(staging, blockCnt, countIn, countHost, listOut, countOut);
}

__global__ void k_flag_cells(const int *__restrict__ tasks, long long n, unsigned char *cellFlag) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        cellFlag[__ldg(tasks + 3 * t)] = 1;
}
void launch_flag_cells(const int *tasks, long long n, unsigned char *cellFlag, cudaStream_t s) {
    if (n <= 0) return;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    ++g_launchCount; k_flag_cells<<<(unsigned)blocks, 256, 0, s>>>(tasks, n, cellFlag);
}

// RefinementsRequired[c] += 1 for every flagged original cell, flags cleared for the next round
// (src/NumericalIntegrator3d.cu:461-499)
__global__ void k_bump(unsigned char *cellFlag, unsigned char *refinements, int nc) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nc && cellFlag[c]) {
        if (refinements) refinements[c] += 1;
        cellFlag[c] = 0;
    }
}
void launch_bump(unsigned char *cellFlag, unsigned char *refinements, int nc, cudaStream_t s) {
    if (nc > 0) { ++g_launchCount; k_bump<<<(nc + 255) / 256, 256, 0, s>>>(cellFlag, refinements, nc); }
}

// ---------------------------------------------------------------------------------------------------------
// closed-form singular integral (adjacent classes) + final assembly, fused
// ---------------------------------------------------------------------------------------------------------
template <int CLS>
__global__ void __launch_bounds__(128)
k_finalize(PackedMesh pm, const double *__restrict__ verts, const int *__restrict__ tasks, long long n, double *bufA,
           const double *__restrict__ bufB, const QueueState *__restrict__ qs, double *__restrict__ results, QueueState *qsMut) {
    if (n <= 0) return;       // (the warm-up launch of preload_kernels)
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < n;
    if (!active) t = n - 1;   // tail lanes redo the last task without storing: the warp votes inside the closed forms stay full-mask
    // result-buffer ping-pong of the reference's adaptive loop (src/evaluators/evaluatorJ3DK.cu:976): after L rounds
    // the live buffer is A for even L, B for odd L; slots whose task converged earlier keep what that buffer last held.
    const double *src = (qs->lastRound & 1) ? bufB : bufA;
    const double2 *sp = reinterpret_cast<const double2 *>(src + 4 * t);
    const double2 v0 = sp[0], v1 = sp[1];
    d4 I = {v0.x, v0.y, v1.x, v1.y};
    const int i = __ldg(tasks + 3 * t), j = __ldg(tasks + 3 * t + 1);
    const int stride = pm.stride;
    const d3 nj = ld3(pm.tri + PK_N * stride, stride, j);
    const double Si = __ldg(pm.tri + PK_S * stride + i);
    if (CLS != 2) {
        const tri3 ti = ldtri(pm.cells, i), tj = ldtri(pm.cells, j);
        int si, sj;
        if (CLS == 0) shifts_vertex(ti, tj, si, sj);
        else shifts_edge(ti, tj, si, sj);
        const tri3 ri = rot_left(ti, si), rj = rot_left(tj, sj);
        const d3 ni = ld3(pm.tri + PK_N * stride, stride, i);
        if (CLS == 0) {
            bool bad = false;
            I = I + integral_singular_vertex(ldv(verts, ri.a), ldv(verts, ri.b), ldv(verts, ri.c), ldv(verts, rj.a), ldv(verts, rj.b),
                                             ldv(verts, rj.c), ni, nj, Si, &bad);
            if (bad && active) {
                printf("Orientation is incorrect for pair (%d, %d)\n", i, j);  // same text as the reference (:701-702)
                atomicAdd(&qsMut->orientationWarnings, 1);
            }
        } else {
            I = I + integral_singular_edge(ldv(verts, ri.a), ldv(verts, ri.b), ldv(verts, ri.c), ldv(verts, rj.a), ldv(verts, rj.b),
                                           ldv(verts, rj.c), ni, nj, Si);
        }
    }
    const d3 J = assemble_J(I, nj, Si, CLS == 0);
    if (!active) return;
    double2 *op = reinterpret_cast<double2 *>(bufA + 4 * t);
    op[0] = make_double2(I.x, I.y);
    op[1] = make_double2(I.z, I.w);
    results[3 * t] = J.x; results[3 * t + 1] = J.y; results[3 * t + 2] = J.z;
}

void launch_finalize(int cls, const PackedMesh &pm, const double *verts, const int *tasks, long long n, double *bufA4, const double *bufB4,
                     const QueueState *qs, double *results3, QueueState *qsMut, cudaStream_t s) {
    if (n <= 0) return;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (cls == 0) { ++g_launchCount; k_finalize<0><<<blocks, 128, 0, s>>>(pm, verts, tasks, n, bufA4, bufB4, qs, results3, qsMut); }
    else if (cls == 1) { ++g_launchCount; k_finalize<1><<<blocks, 128, 0, s>>>(pm, verts, tasks, n, bufA4, bufB4, qs, results3, qsMut); }
    else { ++g_launchCount; k_finalize<2><<<blocks, 128, 0, s>>>(pm, verts, tasks, n, bufA4, bufB4, qs, results3, qsMut); }
}

// ---------------------------------------------------------------------------------------------------------
// boundary helpers: (i,j)/(j,i) defect, reversed pairs, neighbour classification
// ---------------------------------------------------------------------------------------------------------
__global__ void k_symmetry_error(const double *__restrict__ results, long long n, double *errors) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const d3 a = {results[3 * t], results[3 * t + 1], results[3 * t + 2]};
    const d3 b = {results[3 * (n + t)], results[3 * (n + t) + 1], results[3 * (n + t) + 2]};
    const double delta = l1(a + b) / fmax(l1(a), l1(b));
    errors[t] = delta;
    errors[n + t] = delta;
}
void launch_symmetry_error(const double *results3, long long nHalf, double *errors, cudaStream_t s) {
    if (nHalf > 0) { ++g_launchCount; k_symmetry_error<<<(unsigned)((nHalf + 255) / 256), 256, 0, s>>>(results3, nHalf, errors); }
}

// per-class checksum of the results (sum of components and sum of |J|_1): the small 'metric' read back by the
// device-resident end-to-end path
__global__ void __launch_bounds__(256) k_checksum(const double *__restrict__ results, long long n, double *sums4) {
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const double x = results[3 * t], y = results[3 * t + 1], z = results[3 * t + 2];
        a[0] += x; a[1] += y; a[2] += z; a[3] += fabs(x) + fabs(y) + fabs(z);
    }
    for (int k = 0; k < 4; ++k) {
        for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
        if ((threadIdx.x & 31) == 0) atomicAdd(sums4 + k, a[k]);
    }
}
void launch_checksum(const double *results3, long long n, double *sums4, int numSMs, cudaStream_t s) {
    if (n > 0) { ++g_launchCount; k_checksum<<<numSMs * 8, 256, 0, s>>>(results3, n, sums4); }
}

__global__ void k_add_reversed(int *tasks, long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    tasks[3 * (n + t)] = tasks[3 * t + 1];
    tasks[3 * (n + t) + 1] = tasks[3 * t];
    tasks[3 * (n + t) + 2] = (int)(n + t);
}
void launch_add_reversed(int *tasks3, long long n, cudaStream_t s) {
    if (n > 0) { ++g_launchCount; k_add_reversed<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tasks3, n); }
}

// ---------------------------------------------------------------------------------------------------------
// roofline denominators measured on the box: FP64 pipe (DFMA) and XU pipe (MUFU.RSQ64H) peak rates
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_peak_dfma(double *sink, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int k = 0; k < iters; ++k) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;
}
// same, but every DFMA reads three DISTINCT 64-bit register operands (the pattern of real code: dot products, Horner
// steps with register coefficients), which exercises the register-file operand bandwidth as well as the pipe
__global__ void __launch_bounds__(256) k_peak_dfma3(double *sink, const double *seed, int iters) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 1e-9 + k; b[k] = seed[k] + threadIdx.x * 1e-12; c[k] = seed[8 + k] - threadIdx.x * 1e-12; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(b[k], c[(k + 3) & 7], a[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = fma(a[k], c[k], b[(k + 5) & 7]);
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += a[k] + b[k];
    if (r == 123.456) sink[0] = r;
}
// DFMA chains interleaved with NI independent integer instructions per DFMA: does a warp-wide FP64 instruction (two cycles of
// the sub-partition's FP64 pipe) also hold the DISPATCH port for two cycles, or can integer work issue in its shadow?
template <int NI>
__global__ void __launch_bounds__(256) k_peak_mix(double *sink, int iters, int seed) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    const int t = threadIdx.x * 2654435761u;   // per-thread values: keeps the integer work off the uniform datapath
    int i0 = seed ^ t, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    for (int k = 0; k < iters; ++k) {
#define I2_MIX(a, i)                                                           \
        a = fma(a, m, c);                                                      \
        if (NI >= 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i) : "r"(t), "r"(k)); \
        if (NI >= 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i) : "r"(k), "r"(t)); \
        if (NI >= 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i) : "r"(t), "r"(seed));
        I2_MIX(a0, i0) I2_MIX(a1, i1) I2_MIX(a2, i2) I2_MIX(a3, i3) I2_MIX(a4, i4) I2_MIX(a5, i5) I2_MIX(a6, i6) I2_MIX(a7, i7)
#undef I2_MIX
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    const int q = i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
    if (r == 123.456 || q == 0x7fffffff) sink[0] = r + q;
}
__global__ void __launch_bounds__(256) k_peak_mufu(double *sink, int iters) {
    double a0 = 1.5 + threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    for (int k = 0; k < iters; ++k) {
        asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a0));
        asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a1));
        asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a2));
        asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a3));
    }
    const double r = (a0 + a1) + (a2 + a3);
    if (r == 123.456) sink[0] = r;
}
__global__ void k_selftest_math(int op, const double *__restrict__ a, const double *__restrict__ b, long long n, double *out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double r;
    if (op == 0) r = fast_sqrt(a[t]);
    else if (op == 1) r = fast_rcp(a[t]);
    else if (op == 2) r = log_ratio(a[t], b[t]);
    else r = atan2_fast(a[t], b[t]);
    out[t] = r;
}
void launch_selftest_math(int op, const double *a, const double *b, long long n, double *out, cudaStream_t s) {
    if (n > 0) { ++g_launchCount; k_selftest_math<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(op, a, b, n, out); }
}
cudaError_t preload_rest() {
    cudaFuncAttributes a;
    cudaError_t e;
    k_finalize<0><<<1, 128>>>(PackedMesh{nullptr, nullptr, 0, 0}, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    k_finalize<1><<<1, 128>>>(PackedMesh{nullptr, nullptr, 0, 0}, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    k_finalize<2><<<1, 128>>>(PackedMesh{nullptr, nullptr, 0, 0}, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    k_symmetry_error<<<1, 32>>>(nullptr, 0, nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
#define I2_TOUCH(...) do { e = cudaFuncGetAttributes(&a, (const void *)(__VA_ARGS__)); if (e != cudaSuccess) return e; } while (0)
    I2_TOUCH(k_finalize<0>);
    I2_TOUCH(k_finalize<1>);
    I2_TOUCH(k_finalize<2>);
    I2_TOUCH(k_compare);
#undef I2_TOUCH
    return cudaSuccess;
}

void launch_peak_dfma(double *sink, int iters, int blocks, cudaStream_t s) { ++g_launchCount; k_peak_dfma<<<blocks, 256, 0, s>>>(sink, iters); }
void launch_peak_dfma3(double *sink, const double *seed, int iters, int blocks, cudaStream_t s) { ++g_launchCount; k_peak_dfma3<<<blocks, 256, 0, s>>>(sink, seed, iters); }
void launch_peak_mix(int ni, double *sink, int iters, int blocks, cudaStream_t s) {
    ++g_launchCount;
    if (ni <= 0) k_peak_mix<0><<<blocks, 256, 0, s>>>(sink, iters, 12345);
    else if (ni == 1) k_peak_mix<1><<<blocks, 256, 0, s>>>(sink, iters, 12345);
    else if (ni == 2) k_peak_mix<2><<<blocks, 256, 0, s>>>(sink, iters, 12345);
    else k_peak_mix<3><<<blocks, 256, 0, s>>>(sink, iters, 12345);
}
void launch_peak_mufu(double *sink, int iters, int blocks, cudaStream_t s) { ++g_launchCount; k_peak_mufu<<<blocks, 256, 0, s>>>(sink, iters); }

}  // namespace i2
