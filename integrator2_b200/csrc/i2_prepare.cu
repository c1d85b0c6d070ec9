// Neighbour classification by vertex incidence (CSR) and sharded task-list construction.
//
// Replaces kDetermineNeighborType + fillNeightborsLists (src/Mesh3d.cu:93-142, 243-262) and the task-list part of
// Evaluator3D::runAllPairs (src/evaluators/evaluator3d.cu:122-169: copy of the pair lists + kAddReversedPairs).
// The reference compares every pair of triangles (O(N^2)) and appends to three global lists through atomic counters.
// Here:
//   * the two small classes (vertex-/edge-adjacent) come from the vertex -> triangle incidence: the partners of
//     triangle i are the triangles listed under its three vertices; a partner that appears once shares one vertex
//     (simple), twice an edge (attached), three times all vertices (dropped, like the reference: no branch for
//     commonPoints == 3).  O(N * valence^2), no pass over the pair matrix;
//   * the regular class is implicit: row i holds (nc - 1 - i) - partners(i) pairs, so every row's first slot follows
//     from a prefix sum and the list can be materialised for ANY contiguous range of slots independently — each GPU
//     fills only its shard, and the reversed pairs (j, i, n + slot) are written by the same kernel;
//   * all lists come out in lexicographic (i, j) order whatever the scheduling (the reference's order depends on the
//     atomics); slot = rank in that order, as the oracle's lists.
#include "i2_kernels.cuh"

namespace i2 {

static __device__ __forceinline__ tri3 ldcell(const int *__restrict__ cells, int c) {
    return {__ldg(cells + 3 * (long long)c), __ldg(cells + 3 * (long long)c + 1), __ldg(cells + 3 * (long long)c + 2)};
}

// ---- vertex -> triangle incidence --------------------------------------------------------------------------------
__global__ void k_incidence_count(const int *__restrict__ cells, int nc, int *__restrict__ vcount) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < 3 * nc) atomicAdd(vcount + __ldg(cells + e), 1);
}

// exclusive prefix sum of n ints by ONE CTA (n is a vertex or row count: 1e4 .. 1e6), out[n] = total
__global__ void __launch_bounds__(1024) k_scan_int(const int *__restrict__ in, int n, int *__restrict__ out) {
    __shared__ int warpTot[32];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int k = lo; k < hi; ++k) s += in[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = s;
    for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += v; }
    if (lane == 31) warpTot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warpTot[lane];
        for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += v; }
        warpTot[lane] = w;
    }
    __syncthreads();
    int run = inc - s + (warp ? warpTot[warp - 1] : 0);
    for (int k = lo; k < hi; ++k) { const int v = in[k]; out[k] = run; run += v; }
    if (threadIdx.x == 1023) out[n] = warpTot[31];
}

__global__ void k_incidence_fill(const int *__restrict__ cells, int nc, const int *__restrict__ voff, int *__restrict__ cursor,
                                 int *__restrict__ inc) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * nc) return;
    const int v = __ldg(cells + e);
    inc[voff[v] + atomicAdd(cursor + v, 1)] = e / 3;   // order inside a vertex's segment is arbitrary: consumers rank by j
}

// ---- partners of one row from the incidence -----------------------------------------------------------------------
// One warp per row i.  cand[] (shared memory, kCandCap entries per warp) receives every j > i listed under one of i's
// three vertices, once per shared vertex; mult[e] = number of entries equal to cand[e] (1, 2 or 3).
// Rows with more candidates than kCandCap (a vertex of valence > ~60) take the slow path: the warp scans all j > i and
// compares vertex ids, like the reference's kernel does for every row.
constexpr int kCandCap = 192;
constexpr int kRowWarps = 8;

static __device__ __forceinline__ int common_vertices(tri3 a, tri3 b) {
    return (a.a == b.a || a.a == b.b || a.a == b.c) + (a.b == b.a || a.b == b.b || a.b == b.c) + (a.c == b.a || a.c == b.b || a.c == b.c);
}

static __device__ int gather_partners(const int *__restrict__ cells, const int *__restrict__ voff, const int *__restrict__ inc, int i,
                                      int *cand, unsigned char *mult, bool both = false) {
    const int lane = threadIdx.x & 31;
    const tri3 t = ldcell(cells, i);
    const int v[3] = {t.a, t.b, t.c};
    int total = 0;
    for (int k = 0; k < 3; ++k) total += voff[v[k] + 1] - voff[v[k]];
    if (total > kCandCap) return -1;
    int n = 0;
    for (int k = 0; k < 3; ++k) {
        const int lo = voff[v[k]], len = voff[v[k] + 1] - lo;
        for (int base = 0; base < len; base += 32) {
            const int e = base + lane;
            const int j = e < len ? inc[lo + e] : -1;
            const bool keep = both ? (j >= 0 && j != i) : j > i;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) cand[n + __popc(m & ((1u << lane) - 1u))] = j;
            n += __popc(m);
        }
    }
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
        const int j = cand[e];
        int c = 0;
        for (int f = 0; f < n; ++f) c += cand[f] == j;
        mult[e] = (unsigned char)c;
    }
    __syncwarp();
    return n;
}

// per row: number of vertex-adjacent, edge-adjacent and dropped (3 shared ids) partners j > i
__global__ void __launch_bounds__(32 * kRowWarps) k_partners_count(const int *__restrict__ cells, int nc, const int *__restrict__ voff,
                                                                  const int *__restrict__ inc, int *__restrict__ cntS, int *__restrict__ cntA,
                                                                  int *__restrict__ cntD) {
    __shared__ int candSh[kRowWarps][kCandCap];
    __shared__ unsigned char multSh[kRowWarps][kCandCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kRowWarps + warp;
    if (i >= nc) return;
    int *cand = candSh[warp];
    unsigned char *mult = multSh[warp];
    const int n = gather_partners(cells, voff, inc, i, cand, mult);
    int c1 = 0, c2 = 0, c3 = 0;
    if (n >= 0) {
        for (int e = lane; e < n; e += 32) { const int m = mult[e]; c1 += m == 1; c2 += m == 2; c3 += m == 3; }
    } else {
        const tri3 a = ldcell(cells, i);
        for (int j = i + 1 + lane; j < nc; j += 32) { const int c = common_vertices(a, ldcell(cells, j)); c1 += c == 1; c2 += c == 2; c3 += c == 3; }
    }
    c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
    // fast path: an edge-adjacent partner owns two candidate entries, a coincident one three
    if (lane == 0) { cntS[i] = c1; cntA[i] = n >= 0 ? c2 / 2 : c2; cntD[i] = n >= 0 ? c3 / 3 : c3; }
}

// first slot of every row in the three lists (exclusive prefix sums; the regular class in 64 bits) + totals[3]
__global__ void __launch_bounds__(1024) k_row_offsets(const int *__restrict__ cntS, const int *__restrict__ cntA, const int *__restrict__ cntD, int nc,
                                                     unsigned long long *__restrict__ rowOff /* [3][nc + 1] */, unsigned long long *__restrict__ totals) {
    __shared__ unsigned long long warpTot[3][32];
    const int per = (nc + 1023) / 1024;
    const int lo = min(nc, (int)threadIdx.x * per), hi = min(nc, lo + per);
    unsigned long long s[3] = {0, 0, 0};
    for (int i = lo; i < hi; ++i) {
        const int a = cntS[i], b = cntA[i];
        s[0] += a; s[1] += b; s[2] += (unsigned long long)(nc - 1 - i - a - b - cntD[i]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc[3];
    for (int k = 0; k < 3; ++k) {
        unsigned long long v = s[k];
        for (int off = 1; off < 32; off <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += u; }
        inc[k] = v;
        if (lane == 31) warpTot[k][warp] = v;
    }
    __syncthreads();
    if (warp < 3) {
        unsigned long long w = warpTot[warp][lane];
        for (int off = 1; off < 32; off <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += u; }
        warpTot[warp][lane] = w;
    }
    __syncthreads();
    unsigned long long run[3];
    for (int k = 0; k < 3; ++k) run[k] = inc[k] - s[k] + (warp ? warpTot[k][warp - 1] : 0ull);
    for (int i = lo; i < hi; ++i) {
        const int a = cntS[i], b = cntA[i];
        rowOff[i] = run[0]; rowOff[(size_t)(nc + 1) + i] = run[1]; rowOff[2 * (size_t)(nc + 1) + i] = run[2];
        run[0] += a; run[1] += b; run[2] += (unsigned long long)(nc - 1 - i - a - b - cntD[i]);
    }
    if (threadIdx.x == 1023)
        for (int k = 0; k < 3; ++k) { rowOff[k * (size_t)(nc + 1) + nc] = warpTot[k][31]; totals[k] = warpTot[k][31]; }
}

static __device__ __forceinline__ void put_task(int *__restrict__ list, unsigned long long at, int i, int j, long long k) {
    list[3 * at] = i; list[3 * at + 1] = j; list[3 * at + 2] = (int)k;
}

// Adjacent lists, whole (they are small: ~6N and ~1.5N pairs): tasks[p] = (i, j, p), tasks[n + p] = (j, i, n + p) when
// `reversed` (the ordered task list of Evaluator3D::runAllPairs), else only the pairs (Mesh3D's lists).
__global__ void __launch_bounds__(32 * kRowWarps) k_partners_fill(const int *__restrict__ cells, int nc, const int *__restrict__ voff,
                                                                 const int *__restrict__ inc, const unsigned long long *__restrict__ rowOff,
                                                                 int *__restrict__ simple, int *__restrict__ attached, int reversed) {
    __shared__ int candSh[kRowWarps][kCandCap];
    __shared__ unsigned char multSh[kRowWarps][kCandCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kRowWarps + warp;
    if (i >= nc) return;
    int *cand = candSh[warp];
    unsigned char *mult = multSh[warp];
    const unsigned long long offS = rowOff[i], offA = rowOff[(size_t)(nc + 1) + i];
    const unsigned long long nS = rowOff[nc], nA = rowOff[(size_t)(nc + 1) + nc];
    const int n = gather_partners(cells, voff, inc, i, cand, mult);
    if (n >= 0) {
        for (int e = lane; e < n; e += 32) {
            const int j = cand[e], m = mult[e];
            if (m > 2) continue;
            // rank among the distinct partners of the same class with smaller j: every one of them owns exactly m entries.
            // an entry equal to j at a lower index means this one is a duplicate (edge-adjacent partners appear twice).
            int less = 0;
            bool first = true;
            for (int f = 0; f < n; ++f) {
                const int jf = cand[f];
                less += (jf < j) && (mult[f] == m);
                first = first && !(jf == j && f < e);
            }
            if (!first) continue;
            int *list = m == 1 ? simple : attached;
            if (!list) continue;
            const unsigned long long p = (m == 1 ? offS : offA) + (unsigned long long)(less / m), tot = m == 1 ? nS : nA;
            put_task(list, p, i, j, (long long)p);
            if (reversed) put_task(list, tot + p, j, i, (long long)(tot + p));
        }
    } else {
        const tri3 a = ldcell(cells, i);
        unsigned long long runS = offS, runA = offA;
        for (int base = i + 1; base < nc; base += 32) {
            const int j = base + lane;
            const int c = j < nc ? common_vertices(a, ldcell(cells, j)) : 0;
            const unsigned mS = __ballot_sync(0xffffffffu, c == 1), mA = __ballot_sync(0xffffffffu, c == 2);
            const unsigned below = (1u << lane) - 1u;
            if (c == 1 && simple) {
                const unsigned long long p = runS + __popc(mS & below);
                put_task(simple, p, i, j, (long long)p);
                if (reversed) put_task(simple, nS + p, j, i, (long long)(nS + p));
            }
            if (c == 2 && attached) {
                const unsigned long long p = runA + __popc(mA & below);
                put_task(attached, p, i, j, (long long)p);
                if (reversed) put_task(attached, nA + p, j, i, (long long)(nA + p));
            }
            runS += __popc(mS); runA += __popc(mA);
        }
    }
}

// Regular list for the forward slots [fLo, fHi) of the class: out[p - fLo] = (i, j, p) and, when outRev != nullptr,
// outRev[p - fLo] = (j, i, n + p) (n = number of regular pairs of the whole mesh).  One CTA per row, rows dealt
// round-robin to a persistent grid; the rows that intersect the slot range are found by binary search in the row offsets.
constexpr int kFillThreads = 256;
__global__ void __launch_bounds__(kFillThreads) k_regular_fill(const int *__restrict__ cells, int nc, const unsigned long long *__restrict__ rowOffR,
                                                              unsigned long long fLo, unsigned long long fHi, int *__restrict__ out,
                                                              int *__restrict__ outRev) {
    __shared__ int warpCnt[kFillThreads / 32];
    if (fHi <= fLo) return;
    const unsigned long long n = rowOffR[nc];
    // first row whose range ends beyond fLo, last row that starts before fHi
    auto upper = [&](unsigned long long x) {   // number of rows r with rowOffR[r] <= x, minus one = row containing slot x
        int lo = 0, hi = nc;                  // invariant: rowOffR[lo] <= x < rowOffR[hi] (rowOffR[nc] = n > x)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (rowOffR[mid] <= x) lo = mid; else hi = mid; }
        return lo;
    };
    const int rowLo = upper(fLo), rowHi = upper(fHi - 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = rowLo + blockIdx.x; i <= rowHi; i += gridDim.x) {
        const tri3 a = ldcell(cells, i);
        unsigned long long run = rowOffR[i];
        const unsigned long long rowEnd = rowOffR[i + 1];
        if (rowEnd <= fLo || run >= fHi) continue;
        for (int base = i + 1; base < nc && run < fHi; base += kFillThreads) {
            const int j = base + threadIdx.x;
            const bool reg = j < nc && common_vertices(a, ldcell(cells, j)) == 0;
            const unsigned m = __ballot_sync(0xffffffffu, reg);
            if (lane == 0) warpCnt[warp] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < kFillThreads / 32; ++w) { const int c = warpCnt[w]; before += w < warp ? c : 0; total += c; }
            if (reg) {
                const unsigned long long p = run + before + __popc(m & ((1u << lane) - 1u));
                if (p >= fLo && p < fHi) {
                    put_task(out, p - fLo, i, j, (long long)p);
                    if (outRev) put_task(outRev, p - fLo, j, i, (long long)(n + p));
                }
            }
            run += total;
            __syncthreads();
        }
    }
}

// Predicted cost of the regular pairs of every row under automatic error control, in level-0 pair integrations:
// 5 (rounds 0 and 1 are unconditional) + 66.7 * P(rho), rho = |c_i - c_j| / sqrt(max(S_i, S_j)), P = probability of
// surviving the first compare round (measured with the CPU oracle on s5m.dat).  upperOnly: only
// j > i, counted twice (a shard owns a pair in both orders).  Pairs that share a vertex are not excluded: they are
// O(valence) per row and their class is integrated separately.
__global__ void __launch_bounds__(256) k_row_cost(const double *__restrict__ tri, int stride, int nc, int upperOnly, double *__restrict__ cost) {
    __shared__ double red[8];
    const int i = blockIdx.x;
    auto centroid = [&](int t, double &x, double &y, double &z) {
        x = (tri[PK_A * stride + t] + tri[PK_B * stride + t] + tri[PK_C * stride + t]) * (1.0 / 3.0);
        y = (tri[(PK_A + 1) * stride + t] + tri[(PK_B + 1) * stride + t] + tri[(PK_C + 1) * stride + t]) * (1.0 / 3.0);
        z = (tri[(PK_A + 2) * stride + t] + tri[(PK_B + 2) * stride + t] + tri[(PK_C + 2) * stride + t]) * (1.0 / 3.0);
    };
    double xi, yi, zi;
    centroid(i, xi, yi, zi);
    const double Si = tri[PK_S * stride + i];
    double acc = 0.0;
    for (int j = (upperOnly ? i + 1 : 0) + threadIdx.x; j < nc; j += blockDim.x) {
        if (j == i) continue;
        double xj, yj, zj;
        centroid(j, xj, yj, zj);
        const double d2 = (xi - xj) * (xi - xj) + (yi - yj) * (yi - yj) + (zi - zj) * (zi - zj);
        const float rho2 = (float)(d2 / fmax(Si, tri[PK_S * stride + j]));
        // P(rho) over the edges 1, 1.5, 2, 3, 4, 6, 10 (multigpu.py: _RHO_EDGES / _P_UNCONVERGED)
        const float p = rho2 < 1.f ? 0.69f : rho2 < 2.25f ? 0.42f : rho2 < 4.f ? 0.13f : rho2 < 9.f ? 0.031f : rho2 < 16.f ? 0.0066f
                        : rho2 < 36.f ? 0.0020f : rho2 < 100.f ? 0.0006f : 0.f;
        acc += 5.0 + 66.7 * (double)p;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        cost[i] = upperOnly ? 2.0 * s : s;
    }
}

// ---- row-major adjacent lists (the operator apply, i2_apply_*): ALL partners j != i of every row, sorted by (i, j) ----
// counts per row (both sides)
__global__ void __launch_bounds__(32 * kRowWarps) k_partners_count_both(const int *__restrict__ cells, int nc, const int *__restrict__ voff,
                                                                       const int *__restrict__ inc, int *__restrict__ cntS, int *__restrict__ cntA) {
    __shared__ int candSh[kRowWarps][kCandCap];
    __shared__ unsigned char multSh[kRowWarps][kCandCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kRowWarps + warp;
    if (i >= nc) return;
    const int n = gather_partners(cells, voff, inc, i, candSh[warp], multSh[warp], true);
    int c1 = 0, c2 = 0;
    if (n >= 0) {
        for (int e = lane; e < n; e += 32) { const int m = multSh[warp][e]; c1 += m == 1; c2 += m == 2; }
    } else {
        const tri3 a = ldcell(cells, i);
        for (int j = lane; j < nc; j += 32) { const int c = j == i ? 0 : common_vertices(a, ldcell(cells, j)); c1 += c == 1; c2 += 2 * (c == 2); }
    }
    c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2);
    if (lane == 0) { cntS[i] = c1; cntA[i] = c2 / 2; }
}

// tasks (i, j, slot) of rows [rowLo, rowHi): slot = offset of the row inside the block + rank of j among the row's partners
// of the class.  offS / offA = exclusive prefix sums of the both-sided counts over ALL rows.
__global__ void __launch_bounds__(32 * kRowWarps) k_partners_fill_rows(const int *__restrict__ cells, int nc, const int *__restrict__ voff,
                                                                      const int *__restrict__ inc, const int *__restrict__ offS,
                                                                      const int *__restrict__ offA, int rowLo, int rowHi,
                                                                      int *__restrict__ simple, int *__restrict__ attached) {
    __shared__ int candSh[kRowWarps][kCandCap];
    __shared__ unsigned char multSh[kRowWarps][kCandCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = rowLo + blockIdx.x * kRowWarps + warp;
    if (i >= rowHi) return;
    int *cand = candSh[warp];
    unsigned char *mult = multSh[warp];
    const int baseS = offS[i] - offS[rowLo], baseA = offA[i] - offA[rowLo];
    const int n = gather_partners(cells, voff, inc, i, cand, mult, true);
    if (n >= 0) {
        for (int e = lane; e < n; e += 32) {
            const int j = cand[e], m = mult[e];
            if (m > 2) continue;
            int less = 0;
            bool first = true;
            for (int f = 0; f < n; ++f) {
                const int jf = cand[f];
                less += (jf < j) && (mult[f] == m);
                first = first && !(jf == j && f < e);
            }
            if (!first) continue;
            const int p = (m == 1 ? baseS : baseA) + less / m;
            put_task(m == 1 ? simple : attached, (unsigned long long)p, i, j, p);
        }
    } else {
        const tri3 a = ldcell(cells, i);
        int runS = baseS, runA = baseA;
        for (int base = 0; base < nc; base += 32) {
            const int j = base + lane;
            const int c = (j < nc && j != i) ? common_vertices(a, ldcell(cells, j)) : 0;
            const unsigned mS = __ballot_sync(0xffffffffu, c == 1), mA = __ballot_sync(0xffffffffu, c == 2);
            const unsigned below = (1u << lane) - 1u;
            if (c == 1) { const int p = runS + __popc(mS & below); put_task(simple, (unsigned long long)p, i, j, p); }
            if (c == 2) { const int p = runA + __popc(mA & below); put_task(attached, (unsigned long long)p, i, j, p); }
            runS += __popc(mS); runA += __popc(mA);
        }
    }
}

// out[i - rowLo] += sum over the row's tasks of w_j J(K_i, K_j), in list order (deterministic); one thread per row
__global__ void k_row_scatter(const int *__restrict__ tasks, const double *__restrict__ results, const int *__restrict__ off, int rowLo, int rows,
                              const double *__restrict__ weights, double *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int base = off[rowLo];
    double x = 0.0, y = 0.0, z = 0.0;
    for (int t = off[rowLo + r] - base; t < off[rowLo + r + 1] - base; ++t) {
        const double w = weights ? weights[tasks[3 * (long long)t + 1]] : 1.0;
        x = fma(w, results[3 * (long long)t], x); y = fma(w, results[3 * (long long)t + 1], y); z = fma(w, results[3 * (long long)t + 2], z);
    }
    out[3 * r] += x; out[3 * r + 1] += y; out[3 * r + 2] += z;
}

// per-row maximum of the refinement counters' source: refinements[r] = value of the class's per-cell counter for row rowLo + r
__global__ void k_take_rows(const unsigned char *__restrict__ perCell, int rowLo, int rows, unsigned char *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) out[r] = perCell[rowLo + r];
}

__global__ void k_max_vertex_id(const int *__restrict__ cells, int nc, int *maxId) {
    int m = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * nc; e += gridDim.x * blockDim.x) m = max(m, __ldg(cells + e));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(maxId, m);
}

// max and sum of non-negative doubles (NaN-free inputs order like their bit patterns)
__global__ void __launch_bounds__(256) k_error_summary(const double *__restrict__ e, long long n, double *out2) {
    double mx = 0.0, sm = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const double v = e[t];
        mx = fmax(mx, v);
        sm += v;
    }
    for (int off = 16; off > 0; off >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off)); sm += __shfl_xor_sync(0xffffffffu, sm, off); }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(reinterpret_cast<unsigned long long *>(out2), (unsigned long long)__double_as_longlong(mx));
        atomicAdd(out2 + 1, sm);
    }
}

// ---- host-side launchers ---------------------------------------------------------------------------------------------
// scratch layout (ints): vcount[nv + 1] | voff[nv + 1] | cursor[nv] | inc[3 nc] | cntS[nc] | cntA[nc] | cntD[nc]
size_t incidence_scratch_ints(int nv, int nc) { return (size_t)3 * (nv + 1) + (size_t)6 * nc + 8; }

void launch_incidence(const int *cells, int nv, int nc, int *scratch, unsigned long long *rowOff, unsigned long long *totals, cudaStream_t s) {
    if (nc <= 0) return;
    int *vcount = scratch, *voff = vcount + (nv + 1), *cursor = voff + (nv + 1), *inc = cursor + nv, *cntS = inc + 3 * (size_t)nc, *cntA = cntS + nc,
        *cntD = cntA + nc;
    cudaMemsetAsync(vcount, 0, sizeof(int) * (size_t)(nv + 1), s);
    cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)nv, s);
    const int eb = (3 * nc + 255) / 256;
    g_launchCount += 5;
    k_incidence_count<<<eb, 256, 0, s>>>(cells, nc, vcount);
    k_scan_int<<<1, 1024, 0, s>>>(vcount, nv, voff);
    k_incidence_fill<<<eb, 256, 0, s>>>(cells, nc, voff, cursor, inc);
    k_partners_count<<<(nc + kRowWarps - 1) / kRowWarps, 32 * kRowWarps, 0, s>>>(cells, nc, voff, inc, cntS, cntA, cntD);
    k_row_offsets<<<1, 1024, 0, s>>>(cntS, cntA, cntD, nc, rowOff, totals);
}

void launch_partners_fill(const int *cells, int nv, int nc, const int *scratch, const unsigned long long *rowOff, int *simple, int *attached,
                          bool reversed, cudaStream_t s) {
    if (nc <= 0 || (!simple && !attached)) return;
    const int *voff = scratch + (nv + 1), *inc = voff + (nv + 1) + nv;
    ++g_launchCount;
    k_partners_fill<<<(nc + kRowWarps - 1) / kRowWarps, 32 * kRowWarps, 0, s>>>(cells, nc, voff, inc, rowOff, simple, attached, reversed ? 1 : 0);
}

void launch_regular_fill(const int *cells, int nc, const unsigned long long *rowOff, unsigned long long fLo, unsigned long long fHi, int *out,
                         int *outRev, int numSMs, cudaStream_t s) {
    if (nc <= 0 || fHi <= fLo || !out) return;
    ++g_launchCount;
    k_regular_fill<<<numSMs * 8, kFillThreads, 0, s>>>(cells, nc, rowOff + 2 * (size_t)(nc + 1), fLo, fHi, out, outRev);
}

void launch_row_cost(const PackedMesh &pm, bool upperOnly, double *cost, cudaStream_t s) {
    if (pm.nc <= 0) return;
    ++g_launchCount;
    k_row_cost<<<pm.nc, 256, 0, s>>>(pm.tri, pm.stride, pm.nc, upperOnly ? 1 : 0, cost);
}

}  // namespace i2

namespace i2 {
// both-sided per-row counts and their prefix sums: scratch2 = cntS[nc] | cntA[nc] | offS[nc + 1] | offA[nc + 1]
size_t rows_scratch_ints(int nc) { return (size_t)4 * nc + 8; }
void launch_partners_both(const int *cells, int nv, int nc, const int *scratch, int *scratch2, cudaStream_t s) {
    if (nc <= 0) return;
    const int *voff = scratch + (nv + 1), *inc = voff + (nv + 1) + nv;
    int *cntS = scratch2, *cntA = cntS + nc, *offS = cntA + nc, *offA = offS + (nc + 1);
    g_launchCount += 3;
    k_partners_count_both<<<(nc + kRowWarps - 1) / kRowWarps, 32 * kRowWarps, 0, s>>>(cells, nc, voff, inc, cntS, cntA);
    k_scan_int<<<1, 1024, 0, s>>>(cntS, nc, offS);
    k_scan_int<<<1, 1024, 0, s>>>(cntA, nc, offA);
}
void launch_partners_fill_rows(const int *cells, int nv, int nc, const int *scratch, const int *scratch2, int rowLo, int rowHi, int *simple,
                               int *attached, cudaStream_t s) {
    if (rowHi <= rowLo) return;
    const int *voff = scratch + (nv + 1), *inc = voff + (nv + 1) + nv;
    const int *offS = scratch2 + 2 * (size_t)nc, *offA = offS + (nc + 1);
    ++g_launchCount;
    k_partners_fill_rows<<<(rowHi - rowLo + kRowWarps - 1) / kRowWarps, 32 * kRowWarps, 0, s>>>(cells, nc, voff, inc, offS, offA, rowLo, rowHi, simple,
                                                                                            attached);
}
void launch_row_scatter(const int *tasks, const double *results, const int *off, int rowLo, int rows, const double *weights, double *out,
                        cudaStream_t s) {
    if (rows <= 0) return;
    ++g_launchCount;
    k_row_scatter<<<(rows + 127) / 128, 128, 0, s>>>(tasks, results, off, rowLo, rows, weights, out);
}
void launch_take_rows(const unsigned char *perCell, int rowLo, int rows, unsigned char *out, cudaStream_t s) {
    if (rows <= 0) return;
    ++g_launchCount;
    k_take_rows<<<(rows + 255) / 256, 256, 0, s>>>(perCell, rowLo, rows, out);
}
}  // namespace i2

namespace i2 {
void launch_max_vertex_id(const int *cells, int nc, int *maxId, cudaStream_t s) {
    if (nc <= 0) return;
    ++g_launchCount;
    int blocks = (3 * nc + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    k_max_vertex_id<<<blocks, 256, 0, s>>>(cells, nc, maxId);
}
void launch_error_summary(const double *errors, long long n, double *out2, int numSMs, cudaStream_t s) {
    if (n <= 0) return;
    ++g_launchCount;
    k_error_summary<<<numSMs * 4, 256, 0, s>>>(errors, n, out2);
}
}  // namespace i2
