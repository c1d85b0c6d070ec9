// C-ABI entry points of libintegrator2_b200.so (declared in include/i2_abi.h).
// Host-side orchestration only: every numerical step is a kernel in i2_kernels.cu; there is no CPU fallback.
#include "i2_context.h"

#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace i2;

#define I2_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return (int)e__;        \
    } while (0)

namespace {

template <class T>
int ensure(T **p, size_t *cap, size_t need) {
    if (need <= *cap && *p) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    if (need == 0) return 0;
    I2_CUDA(cudaMalloc((void **)p, need * sizeof(T)));
    *cap = need;
    return 0;
}

void freeHostState(i2_context *c) {
    auto fr = [](auto *&p, size_t &cap) { if (p) cudaFree(p); p = nullptr; cap = 0; };
    fr(c->hVerts, c->capVerts); fr(c->hNormals, c->capNormals); fr(c->hMeasures, c->capMeasures); fr(c->hCells, c->capCells);
    for (int k = 0; k < 3; ++k) {
        fr(c->hTasks[k], c->capTasks[k]); fr(c->hIntegrals[k], c->capIntegrals[k]); fr(c->hResults[k], c->capResults[k]);
        fr(c->hErrors[k], c->capErrors[k]);
        c->hRefinements[k] = nullptr;
        c->hCount[k] = c->hN[k] = c->hHalf[k] = c->hLo[k] = 0;
    }
    fr(c->hRefAll, c->capRefAll);
    for (int k = 0; k < 2; ++k) fr(c->adjFull[k], c->capAdjFull[k]);
    fr(c->incScratch, c->incCap); fr(c->rowOff, c->rowOffCap); fr(c->rowCost, c->rowCostCap);
    c->hPrepared = false;
}

// Column chunks of the list-free kernels: a FIXED width, so that the order in which a row's partial sums are formed and added
// does not depend on how many rows a call (or a GPU of a multi-GPU run) was given — row sums are bitwise partition-invariant
// when the row blocks start at multiples of 32.  Small meshes get narrower chunks to fill the GPU.
int apply_chunks(int nc) {
    const int width = nc >= (1 << 16) ? 2048 : (nc >= (1 << 13) ? 512 : 256);
    return (nc + width - 1) / width;
}

PackedMesh packed(const i2_context *c) {
    PackedMesh pm;
    pm.tri = c->tri;
    pm.cells = c->cells;
    pm.nc = c->nc;
    pm.stride = c->stride;
    return pm;
}

}  // namespace

extern "C" {

const char *i2_error_string(int code) {
    switch (code) {
    case 0: return "success";
    case I2_E_BADARG: return "i2: bad argument";
    case I2_E_NOMESH: return "i2: no mesh set (i2_set_mesh)";
    case I2_E_NOQUAD: return "i2: no quadrature rule set (i2_set_quadrature)";
    case I2_E_LEVEL: return "i2: refinement level out of range";
    case I2_E_TOOBIG: return "i2: count exceeds 32-bit task slots";
    case I2_E_NCCL: return "i2: NCCL unavailable or failed (multi-GPU layer)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "i2: unknown error";
    }
}

int i2_destroy(i2_context *c);

int i2_create(i2_context **out, int device) {
    if (!out) return I2_E_BADARG;
    *out = nullptr;
    I2_CUDA(cudaSetDevice(device));
    i2_context *c = new i2_context;
    c->device = device;
    c->ownStream = true;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) c->numSMs = prop.multiProcessorCount;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->chunkDone[k], cudaEventDisableTiming);
    for (int k = 0; k < 3 && e == cudaSuccess; ++k) e = cudaMalloc((void **)&c->scr[k].qs, sizeof(QueueState));
    // The side streams carry the two adjacent classes: short chains of small dependent kernels next to the regular class's big grids.
    // At equal priority the block scheduler serves the grid that was launched first, so a small kernel waits until the big one has
    // dispatched ALL its CTAs and the chains crawl behind the regular class instead of hiding under it; at a higher priority
    // their CTAs take the next SM slot that frees (env I2_SIDE_PRIORITY=0 restores equal priorities: A/B knob).
    int prioLeast = 0, prioGreatest = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest);
    const char *pe = getenv("I2_SIDE_PRIORITY");
    const int sidePrio = (pe && atoi(pe) == 0) ? prioLeast : prioGreatest;
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaStreamCreateWithPriority(&c->side[k], cudaStreamNonBlocking, sidePrio);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->sideDone[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->forkEv, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->sumScratch, sizeof(double) * 16);
    if (e == cudaSuccess) e = upload_math_tables(c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = preload_kernels();
    if (e != cudaSuccess) {
        i2_destroy(c);   // releases whatever was created so far
        return (int)e;
    }
    *out = c;
    return 0;
}

int i2_destroy(i2_context *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    freeHostState(c);
    if (c->tri) cudaFree(c->tri);
    for (int k = 0; k < 2; ++k) {
        if (c->side[k]) { cudaStreamSynchronize(c->side[k]); cudaStreamDestroy(c->side[k]); }
        if (c->sideDone[k]) cudaEventDestroy(c->sideDone[k]);
        if (c->chunkDone[k]) cudaEventDestroy(c->chunkDone[k]);
    }
    if (c->forkEv) cudaEventDestroy(c->forkEv);
    if (c->sumScratch) cudaFree(c->sumScratch);
    for (auto &sc : c->scr) {
        if (sc.bufB) cudaFree(sc.bufB);
        for (int k = 0; k < 2; ++k) if (sc.rest[k]) cudaFree(sc.rest[k]);
        if (sc.cellFlag) cudaFree(sc.cellFlag);
        if (sc.blockCnt) cudaFree(sc.blockCnt);
        if (sc.qs) cudaFree(sc.qs);
    }
    for (int k = 0; k < 3; ++k) if (c->prof[k]) cudaEventDestroy(c->prof[k]);
    if (c->rowCounts) cudaFree(c->rowCounts);
    if (c->clsScratch) cudaFree(c->clsScratch);
    if (c->partial) cudaFree(c->partial);
    if (c->depthBuf) cudaFree(c->depthBuf);
    if (c->roundsGraph) cudaGraphExecDestroy(c->roundsGraph);
    if (c->ap.scratch2) cudaFree(c->ap.scratch2);
    if (c->ap.refCells) cudaFree(c->ap.refCells);
    if (c->ap.regular) cudaFree(c->ap.regular);
    for (int k = 0; k < 2; ++k) {
        if (c->ap.tasks[k]) cudaFree(c->ap.tasks[k]);
        if (c->ap.integrals[k]) cudaFree(c->ap.integrals[k]);
        if (c->ap.results[k]) cudaFree(c->ap.results[k]);
    }
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    delete c;
    return 0;
}

int i2_set_stream(i2_context *c, void *s) {
    if (!c) return I2_E_BADARG;
    if (c->ownStream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)s;
    c->ownStream = false;
    return 0;
}

int i2_synchronize(i2_context *c) {
    if (!c) return I2_E_BADARG;
    I2_CUDA(cudaStreamSynchronize(c->stream));
    I2_CUDA(cudaStreamSynchronize(c->copyStream));
    return 0;
}

int i2_set_math_mode(i2_context *c, int mode) {
    if (!c || mode < I2_MATH_STRICT || mode > I2_MATH_FAST_POINTWISE) return I2_E_BADARG;
    c->mathMode = mode;
    return 0;
}

int i2_set_quadrature(i2_context *c, const double *xy, const double *w, int n, int order) {
    if (!c || !xy || !w || n < 1 || n > MAX_GAUSS_POINTS || order < 0 || order > 30) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    double packed4[MAX_GAUSS_POINTS * 4];
    for (int g = 0; g < n; ++g) {
        packed4[4 * g] = xy[2 * g];
        packed4[4 * g + 1] = xy[2 * g + 1];
        packed4[4 * g + 2] = 1.0 - xy[2 * g] - xy[2 * g + 1];  // as NumericalIntegrator3D's ctor (src/NumericalIntegrator3d.cu:202-206)
        packed4[4 * g + 3] = w[g];
    }
    I2_CUDA(upload_quadrature(packed4, n, (double)(1 << order), c->stream, &c->ruleShape13));
    I2_CUDA(cudaStreamSynchronize(c->stream));  // packed4 lives on this stack frame
    c->haveQuad = true;
    return 0;
}

int i2_mesh_geometry(i2_context *c, const double *verts, int nv, const int *cells, int nc, double *normals, double *centers, double *measures) {
    if (!c || nv < 0 || nc < 0 || (nc > 0 && (!verts || !cells))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    launch_geometry(verts, cells, nc, normals, centers, measures, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

int i2_set_mesh(i2_context *c, const double *verts, int nv, const int *cells, int nc, const double *normals, const double *measures) {
    if (!c || nv < 0 || nc < 0 || (nc > 0 && (!verts || !cells || !normals || !measures))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    c->verts = verts; c->cells = cells; c->nv = nv; c->nc = nc;
    c->incidenceValid = false;
    c->ap.prepared = false;
    c->stride = (nc + 31) & ~31;  // keep every component row 256-byte aligned
    int rc = ensure(&c->tri, &c->triCap, (size_t)PK_COUNT * (size_t)c->stride);
    if (rc) return rc;
    launch_pack(verts, cells, normals, measures, nc, c->stride, c->tri, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

int i2_refine_mesh_once(i2_context *c, const double *vin, int nvIn, const int *cin, int ncIn, const double *min, double *vout, int *cout,
                        double *mout) {
    if (!c || nvIn < 0 || ncIn < 0 || (ncIn > 0 && (!vin || !cin || !min || !vout || !cout || !mout))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    if (nvIn > 0) I2_CUDA(cudaMemcpyAsync(vout, vin, sizeof(double) * 3 * nvIn, cudaMemcpyDeviceToDevice, c->stream));
    launch_split_uniform(vin, nvIn, cin, ncIn, min, vout, cout, mout, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

// Classification of a device-resident mesh for Mesh3D: by vertex incidence (see i2_prepare.cu), count then fill so that the
// caller can size its lists exactly.  The number of vertices is 1 + the largest id in d_cells.
int i2_classify_count(i2_context *c, const int *cells, int nc, long long counts[3]) {
    if (!c || !counts || nc < 0 || (nc > 0 && !cells)) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    counts[0] = counts[1] = counts[2] = 0;
    c->rowCountsNc = -1;
    if (nc == 0) return 0;
    cudaStream_t s = c->stream;
    int rc = ensure(&c->rowCounts, &c->rowCap, (size_t)3 * (nc + 1) + 4);
    if (rc) return rc;
    int *maxId = reinterpret_cast<int *>(c->rowCounts + (size_t)3 * (nc + 1) + 3);
    I2_CUDA(cudaMemsetAsync(maxId, 0, sizeof(int), s));
    launch_max_vertex_id(cells, nc, maxId, s);
    int nv = 0;
    I2_CUDA(cudaMemcpyAsync(&nv, maxId, sizeof(int), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaStreamSynchronize(s));
    nv += 1;
    rc = ensure(&c->clsScratch, &c->clsScratchCap, incidence_scratch_ints(nv, nc));
    if (rc) return rc;
    unsigned long long *totalsDev = c->rowCounts + (size_t)3 * (nc + 1);
    launch_incidence(cells, nv, nc, c->clsScratch, c->rowCounts, totalsDev, s);
    I2_CUDA(cudaGetLastError());
    unsigned long long totals[3];
    I2_CUDA(cudaMemcpyAsync(totals, totalsDev, sizeof(totals), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k) counts[k] = (long long)totals[k];
    c->rowCountsNc = nc;
    c->rowCountsNv = nv;
    return 0;
}

int i2_classify_fill(i2_context *c, const int *cells, int nc, int *simple, int *attached, int *notn) {
    if (!c || nc < 0 || (nc > 0 && (!cells || !c->rowCounts))) return I2_E_BADARG;
    if (nc > 0 && c->rowCountsNc != nc) return I2_E_BADARG;   // the offsets must come from i2_classify_count of the same mesh
    I2_CUDA(cudaSetDevice(c->device));
    if (nc == 0) return 0;
    launch_partners_fill(cells, c->rowCountsNv, nc, c->clsScratch, c->rowCounts, simple, attached, false, c->stream);
    if (notn) {
        unsigned long long total = 0;
        I2_CUDA(cudaMemcpyAsync(&total, c->rowCounts + (size_t)3 * (nc + 1) + 2, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
        I2_CUDA(cudaStreamSynchronize(c->stream));
        launch_regular_fill(cells, nc, c->rowCounts, 0ull, total, notn, nullptr, c->numSMs, c->stream);
    }
    I2_CUDA(cudaGetLastError());
    return 0;
}

int i2_add_reversed_pairs(i2_context *c, int *tasks, long long n) {
    if (!c || n < 0 || (n > 0 && !tasks)) return I2_E_BADARG;
    if (2 * n > INT_MAX) return I2_E_TOOBIG;
    I2_CUDA(cudaSetDevice(c->device));
    launch_add_reversed(tasks, n, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

namespace {

// argument checks shared by i2_integrate_class / i2_integrate_all (n == 0 is valid and does nothing)
int check_class_args(const i2_context *c, int cls, const int *tasks, long long n, int level, const double *integrals, const double *results) {
    if (cls < 0 || cls > 2 || n < 0) return I2_E_BADARG;
    if (n == 0) return 0;
    if (!tasks || !integrals || !results) return I2_E_BADARG;
    if (!c->tri) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (level > 12) return I2_E_LEVEL;
    if (n > INT_MAX) return I2_E_TOOBIG;
    return 0;
}

// integration rounds of one class on stream s (no synchronisation); uses the class's own scratch.  At a fixed level with the
// grouped regular kernel the final assembly is fused (*fusedOut = true: the class is complete); otherwise enqueue_finalize
// must follow.  half > 0: the list is [pairs ; reversed pairs] with `half` pairs (see k_regular_grouped).
int enqueue_rounds(i2_context *c, int cls, const int *tasks, long long n, long long half, int level, double *integrals, double *results,
                   unsigned char *refinements, unsigned char *converged, cudaStream_t s, bool profile, bool *fusedOut, int kernelFlags = 0) {
    i2_context::ClassScratch &sc = c->scr[cls];
    const PackedMesh pm = packed(c);
    I2_CUDA(cudaMemsetAsync(sc.qs, 0, sizeof(QueueState), s));
    *fusedOut = false;

    if (level >= 0) {
        if (profile) I2_CUDA(cudaEventRecord(c->prof[0], s));
        // regular pairs with the grouped kernel: the final assembly is fused into the integrate kernel
        *fusedOut = (cls == 2 && c->mathMode == I2_MATH_FAST);
        launch_integrate(cls, c->mathMode, pm, tasks, nullptr, nullptr, n, half, level, integrals, *fusedOut ? results : nullptr, c->numSMs, s,
                         kernelFlags | (c->ruleShape13 ? 2 : 0));
        if (profile) I2_CUDA(cudaEventRecord(c->prof[1], s));
    } else {
        int rc = ensure(&sc.bufB, &sc.bufBCap, (size_t)4 * n);
        if (rc) return rc;
        if ((size_t)n > sc.restCap) {
            size_t cap0 = sc.restCap, cap1 = sc.restCap;
            rc = ensure(&sc.rest[0], &cap0, (size_t)n);
            if (rc) return rc;
            rc = ensure(&sc.rest[1], &cap1, (size_t)n);
            if (rc) return rc;
            sc.restCap = (size_t)n;
        }
        rc = ensure(&sc.cellFlag, &sc.cellFlagCap, (size_t)c->nc);
        if (!rc) rc = ensure(&sc.blockCnt, &sc.blockCntCap, (size_t)kCompareMaxBlocks);
        if (rc) return rc;
        I2_CUDA(cudaMemsetAsync(sc.cellFlag, 0, c->nc, s));

        // round 0: every task on the original control panel; every control panel present in the list is marked
        launch_integrate(cls, c->mathMode, pm, tasks, nullptr, nullptr, n, half, 0, integrals, nullptr, c->numSMs, s, c->ruleShape13 ? 2 : 0);
        launch_flag_cells(tasks, n, sc.cellFlag, s);
        launch_bump(sc.cellFlag, refinements, c->nc, s);
        // rounds 1..5 are enqueued unconditionally; a round whose device-side task count is 0 does nothing.
        for (int m = 1; m <= MAX_REFINE_LEVEL; ++m) {
            double *cur = (m & 1) ? sc.bufB : integrals;
            const double *prev = (m & 1) ? integrals : sc.bufB;
            const int *listIn = m == 1 ? nullptr : sc.rest[0];
            const int *countIn = m == 1 ? nullptr : &sc.qs->count[m - 1];
            launch_integrate(cls, c->mathMode, pm, tasks, listIn, countIn, n, m == 1 ? half : 0, m, cur, nullptr, c->numSMs, s, c->ruleShape13 ? 2 : 0);
            launch_compare(cur, prev, tasks, listIn, countIn, n, sc.rest[1], sc.blockCnt, sc.rest[0], &sc.qs->count[m], sc.cellFlag, converged, sc.qs, m,
                           c->numSMs, s);
            launch_bump(sc.cellFlag, refinements, c->nc, s);
        }
    }
    I2_CUDA(cudaGetLastError());
    return 0;
}

// closed-form singular parts + final assembly; the result buffer of the adaptive ping-pong is selected by the class's
// QueueState::lastRound on the device (a multi-GPU caller overwrites it with the maximum over the ranks first)
int enqueue_finalize(i2_context *c, int cls, const int *tasks, long long n, double *integrals, double *results, cudaStream_t s) {
    i2_context::ClassScratch &sc = c->scr[cls];
    launch_finalize(cls, packed(c), c->verts, tasks, n, integrals, sc.bufB, sc.qs, results, sc.qs, s);
    I2_CUDA(cudaGetLastError());
    return 0;
}

int enqueue_class(i2_context *c, int cls, const int *tasks, long long n, int level, double *integrals, double *results,
                  unsigned char *refinements, unsigned char *converged, cudaStream_t s, bool profile, long long half = 0) {
    bool fused = false;
    int rc = enqueue_rounds(c, cls, tasks, n, half, level, integrals, results, refinements, converged, s, profile, &fused);
    if (!rc && !fused) rc = enqueue_finalize(c, cls, tasks, n, integrals, results, s);
    if (rc) return rc;
    if (profile && level >= 0) I2_CUDA(cudaEventRecord(c->prof[2], s));
    return 0;
}

void fill_stats(const QueueState &h, long long n, int level, i2_stats *stats) {
    stats->last_round = h.lastRound;
    stats->orientation_warnings = h.orientationWarnings;
    stats->integrated[0] = n << (level > 0 ? 2 * level : 0);
    long long before = n;
    for (int m = 1; m <= h.lastRound && m <= MAX_REFINE_LEVEL; ++m) {
        stats->integrated[m] = before << (2 * m);
        stats->unconverged[m] = h.count[m];
        before = h.count[m];
    }
}

}  // namespace

namespace {
int integrate_class_impl(i2_context *c, int cls, const int *tasks, long long n, long long half, int level, double *integrals, double *results,
                         unsigned char *refinements, unsigned char *converged, i2_stats *stats);
}

int i2_integrate_class(i2_context *c, int cls, const int *tasks, long long n, int level, double *integrals, double *results,
                       unsigned char *refinements, unsigned char *converged, i2_stats *stats) {
    return integrate_class_impl(c, cls, tasks, n, 0, level, integrals, results, refinements, converged, stats);
}

int i2_integrate_pairs(i2_context *c, int cls, const int *tasks, long long nHalf, int level, double *integrals, double *results,
                       unsigned char *refinements, unsigned char *converged, i2_stats *stats) {
    if (nHalf < 0) return I2_E_BADARG;
    return integrate_class_impl(c, cls, tasks, 2 * nHalf, nHalf, level, integrals, results, refinements, converged, stats);
}

namespace {
int integrate_class_impl(i2_context *c, int cls, const int *tasks, long long n, long long half, int level, double *integrals, double *results,
                         unsigned char *refinements, unsigned char *converged, i2_stats *stats) {
    if (!c) return I2_E_BADARG;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    int rc = check_class_args(c, cls, tasks, n, level, integrals, results);
    if (rc || n == 0) return rc;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    rc = enqueue_class(c, cls, tasks, n, level, integrals, results, refinements, converged, s, c->profiling, half);
    if (rc) return rc;
    if (stats) {
        QueueState h;
        I2_CUDA(cudaMemcpyAsync(&h, c->scr[cls].qs, sizeof(h), cudaMemcpyDeviceToHost, s));
        I2_CUDA(cudaStreamSynchronize(s));
        fill_stats(h, n, level, stats);
    }
    return 0;
}
}  // namespace

int i2_integrate_all(i2_context *c, const int *const tasks[3], const long long n[3], int level, double *const integrals[3],
                     double *const results[3], unsigned char *const refinements[3], unsigned char *const converged[3], i2_stats stats[3]) {
    if (!c || !tasks || !n || !integrals || !results) return I2_E_BADARG;
    if (stats) std::memset(stats, 0, 3 * sizeof(i2_stats));
    for (int k = 0; k < 3; ++k) {
        const int rc = check_class_args(c, k, tasks[k], n[k], level, integrals[k], results[k]);
        if (rc) return rc;
    }
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // fork: the two adjacent classes (small, latency-bound chains of kernels) go to the side streams and overlap the
    // regular class on the context's stream; join: the context's stream waits for both
    I2_CUDA(cudaEventRecord(c->forkEv, s));
    for (int k = 0; k < 2; ++k) {
        if (n[k] == 0) continue;
        I2_CUDA(cudaStreamWaitEvent(c->side[k], c->forkEv, 0));
        const int rc = enqueue_class(c, k, tasks[k], n[k], level, integrals[k], results[k], refinements ? refinements[k] : nullptr,
                                     converged ? converged[k] : nullptr, c->side[k], false);
        if (rc) return rc;
        I2_CUDA(cudaEventRecord(c->sideDone[k], c->side[k]));
    }
    if (n[2] > 0) {
        const int rc = enqueue_class(c, 2, tasks[2], n[2], level, integrals[2], results[2], refinements ? refinements[2] : nullptr,
                                     converged ? converged[2] : nullptr, s, c->profiling);
        if (rc) return rc;
    }
    for (int k = 0; k < 2; ++k)
        if (n[k] > 0) I2_CUDA(cudaStreamWaitEvent(s, c->sideDone[k], 0));
    if (stats) {
        QueueState h[3];
        for (int k = 0; k < 3; ++k)
            if (n[k] > 0) I2_CUDA(cudaMemcpyAsync(&h[k], c->scr[k].qs, sizeof(QueueState), cudaMemcpyDeviceToHost, s));
        I2_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < 3; ++k)
            if (n[k] > 0) fill_stats(h[k], n[k], level, &stats[k]);
    }
    return 0;
}

int i2_apply_regular(i2_context *c, int rowLo, int rowHi, const double *weights, double *out) {
    if (!c || rowLo < 0 || rowHi < rowLo) return I2_E_BADARG;
    if (!c->tri) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (rowHi > c->nc) return I2_E_BADARG;
    if (rowHi == rowLo) return 0;
    if (!out) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    const int rows = rowHi - rowLo;
    const int chunks = apply_chunks(c->nc);
    int rc = ensure(&c->partial, &c->partialCap, (size_t)chunks * rows * 3);
    if (rc) return rc;
    launch_apply_regular(packed(c), rowLo, rowHi, 0, c->nc, chunks, weights, c->partial, out, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

int i2_apply_regular_adaptive(i2_context *c, int rowLo, int rowHi, const double *weights, double *out, double *outOther,
                              unsigned char *refinements, i2_stats *stats) {
    if (!c || rowLo < 0 || rowHi < rowLo) return I2_E_BADARG;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (!c->tri) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (rowHi > c->nc) return I2_E_BADARG;
    if (rowHi == rowLo) return 0;
    if (!out) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    const int rows = rowHi - rowLo;
    const int chunks = apply_chunks(c->nc);   // 4 lanes per row -> 32 rows per CTA, grid = row blocks x column chunks
    // scratch: 8 doubles of header (6 x 64-bit round counters, the class's last round) + per-chunk partial sums
    int rc = ensure(&c->partial, &c->partialCap, (size_t)8 + (size_t)chunks * rows * 6);
    if (!rc) rc = ensure(&c->depthBuf, &c->depthCap, (size_t)chunks * rows);
    if (rc) return rc;
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(c->partial);
    int *lastRound = reinterpret_cast<int *>(c->partial + 6);
    I2_CUDA(cudaMemsetAsync(c->partial, 0, 8 * sizeof(double), c->stream));
    launch_apply_regular_adaptive(packed(c), rowLo, rowHi, 0, c->nc, chunks, weights, c->partial + 8, c->depthBuf, lastRound, counts, out,
                                  outOther, refinements, c->stream);
    I2_CUDA(cudaGetLastError());
    if (stats) {
        unsigned long long h[8];
        I2_CUDA(cudaMemcpyAsync(h, c->partial, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        I2_CUDA(cudaStreamSynchronize(c->stream));
        int L;
        std::memcpy(&L, &h[6], sizeof(int));
        stats->last_round = L;
        stats->integrated[0] = (long long)h[0];
        long long before = (long long)h[0];
        for (int m = 1; m <= L && m <= MAX_REFINE_LEVEL; ++m) {
            stats->integrated[m] = before << (2 * m);
            stats->unconverged[m] = (long long)h[m];
            before = (long long)h[m];
        }
    }
    return 0;
}

extern "C++" int i2::ensure_incidence(i2_context *c) {
    if (c->incidenceValid) return 0;
    if (!c->cells || c->nc <= 0 || c->nv <= 0) return I2_E_NOMESH;
    int rc = ensure(&c->incScratch, &c->incCap, incidence_scratch_ints(c->nv, c->nc));
    if (!rc) rc = ensure(&c->rowOff, &c->rowOffCap, (size_t)3 * (c->nc + 1) + 3);
    if (rc) return rc;
    launch_incidence(c->cells, c->nv, c->nc, c->incScratch, c->rowOff, c->rowOff + (size_t)3 * (c->nc + 1), c->stream);
    I2_CUDA(cudaGetLastError());
    c->incidenceValid = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// the whole operator for a block of rows: out[i] = sum_j w_j J(K_i, K_j) over ALL neighbour classes; the regular class
// list-free (i2_apply_regular[_adaptive]), the two adjacent classes from row-major lists built by vertex incidence
// ---------------------------------------------------------------------------------------------------------
int i2_apply_prepare(i2_context *c, int rowLo, int rowHi) {
    if (!c || rowLo < 0 || rowHi < rowLo) return I2_E_BADARG;
    if (!c->tri) return I2_E_NOMESH;
    if (rowHi > c->nc) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    i2_context::ApplyState &ap = c->ap;
    ap.prepared = false;
    int rc = i2::ensure_incidence(c);
    if (!rc) rc = ensure(&ap.scratch2, &ap.scratch2Cap, rows_scratch_ints(c->nc));
    if (!rc) rc = ensure(&ap.refCells, &ap.refCellsCap, (size_t)2 * c->nc);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    launch_partners_both(c->cells, c->nv, c->nc, c->incScratch, ap.scratch2, s);
    I2_CUDA(cudaGetLastError());
    const int *offS = ap.scratch2 + 2 * (size_t)c->nc, *offA = offS + (c->nc + 1);
    int h[4];
    I2_CUDA(cudaMemcpyAsync(&h[0], offS + rowLo, sizeof(int), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaMemcpyAsync(&h[1], offS + rowHi, sizeof(int), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaMemcpyAsync(&h[2], offA + rowLo, sizeof(int), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaMemcpyAsync(&h[3], offA + rowHi, sizeof(int), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaStreamSynchronize(s));
    ap.n[0] = h[1] - h[0];
    ap.n[1] = h[3] - h[2];
    for (int k = 0; k < 2; ++k) {
        rc = ensure(&ap.tasks[k], &ap.capTasks[k], (size_t)3 * ap.n[k]);
        if (!rc) rc = ensure(&ap.integrals[k], &ap.capIntegrals[k], (size_t)4 * ap.n[k]);
        if (!rc) rc = ensure(&ap.results[k], &ap.capResults[k], (size_t)3 * ap.n[k]);
        if (rc) return rc;
    }
    launch_partners_fill_rows(c->cells, c->nv, c->nc, c->incScratch, ap.scratch2, rowLo, rowHi, ap.tasks[0], ap.tasks[1], s);
    I2_CUDA(cudaGetLastError());
    ap.rowLo = rowLo;
    ap.rowHi = rowHi;
    ap.prepared = true;
    return 0;
}

int i2_apply_rounds(i2_context *c, int level, const double *weights) {
    if (!c) return I2_E_BADARG;
    if (!c->tri || !c->ap.prepared) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (level > 0) return I2_E_LEVEL;   // the list-free regular kernel exists at level 0 and under error control
    I2_CUDA(cudaSetDevice(c->device));
    i2_context::ApplyState &ap = c->ap;
    cudaStream_t s = c->stream;
    const int rows = ap.rowHi - ap.rowLo;
    if (rows == 0) return 0;
    if (level < 0) I2_CUDA(cudaMemsetAsync(ap.refCells, 0, (size_t)2 * c->nc, s));
    // the two adjacent classes of the block on the side streams, next to the regular class
    I2_CUDA(cudaEventRecord(c->forkEv, s));
    for (int k = 0; k < 2; ++k) {
        cudaStream_t st = c->side[k];
        I2_CUDA(cudaStreamWaitEvent(st, c->forkEv, 0));
        if (ap.n[k] > 0) {
            bool fused = false;
            const int rc = enqueue_rounds(c, k, ap.tasks[k], ap.n[k], 0, level, ap.integrals[k], ap.results[k],
                                          level < 0 ? ap.refCells + (size_t)k * c->nc : nullptr, nullptr, st, false, &fused);
            if (rc) return rc;
        }
        I2_CUDA(cudaEventRecord(c->sideDone[k], st));
    }
    if (level < 0) {
        const int chunks = apply_chunks(c->nc);
        int rc = ensure(&c->partial, &c->partialCap, (size_t)8 + (size_t)chunks * rows * 6);
        if (!rc) rc = ensure(&c->depthBuf, &c->depthCap, (size_t)chunks * rows);
        if (rc) return rc;
        ap.chunks = chunks;
        I2_CUDA(cudaMemsetAsync(c->partial, 0, 8 * sizeof(double), s));
        launch_apply_regular_adaptive(packed(c), ap.rowLo, ap.rowHi, 0, c->nc, chunks, weights, c->partial + 8, c->depthBuf,
                                      reinterpret_cast<int *>(c->partial + 6), reinterpret_cast<unsigned long long *>(c->partial), nullptr, nullptr,
                                      nullptr, s);
    } else {
        const int chunks = apply_chunks(c->nc);
        int rc = ensure(&c->partial, &c->partialCap, (size_t)chunks * rows * 3);
        if (!rc) rc = ensure(&ap.regular, &ap.regularCap, (size_t)rows * 3);
        if (rc) return rc;
        launch_apply_regular(packed(c), ap.rowLo, ap.rowHi, 0, c->nc, chunks, weights, c->partial, ap.regular, s);
    }
    I2_CUDA(cudaGetLastError());
    for (int k = 0; k < 2; ++k) I2_CUDA(cudaStreamWaitEvent(s, c->sideDone[k], 0));
    return 0;
}

int i2_apply_last_rounds(i2_context *c, int last[3], int set) {
    if (!c || !last) return I2_E_BADARG;
    if (!c->ap.prepared || !c->partial) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    int *dev[3] = {&c->scr[0].qs->lastRound, &c->scr[1].qs->lastRound, reinterpret_cast<int *>(c->partial + 6)};
    for (int k = 0; k < 3; ++k) {
        if (set) I2_CUDA(cudaMemcpyAsync(dev[k], &last[k], sizeof(int), cudaMemcpyHostToDevice, c->stream));
        else I2_CUDA(cudaMemcpyAsync(&last[k], dev[k], sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int i2_apply_finish(i2_context *c, int level, const double *weights, double *out, unsigned char *refinements, i2_stats stats[3]) {
    if (!c) return I2_E_BADARG;
    if (stats) std::memset(stats, 0, 3 * sizeof(i2_stats));
    if (!c->tri || !c->ap.prepared) return I2_E_NOMESH;
    if (level > 0) return I2_E_LEVEL;
    I2_CUDA(cudaSetDevice(c->device));
    i2_context::ApplyState &ap = c->ap;
    cudaStream_t s = c->stream;
    const int rows = ap.rowHi - ap.rowLo;
    if (rows == 0) return 0;
    if (!out) return I2_E_BADARG;
    if (level < 0)
        launch_reduce_partials_adaptive(c->partial + 8, c->depthBuf, rows, ap.chunks, reinterpret_cast<int *>(c->partial + 6), out, nullptr,
                                        refinements ? refinements + 2 * (size_t)rows : nullptr, s);
    else
        I2_CUDA(cudaMemcpyAsync(out, ap.regular, sizeof(double) * 3 * rows, cudaMemcpyDeviceToDevice, s));
    const int *offS = ap.scratch2 + 2 * (size_t)c->nc, *offA = offS + (c->nc + 1);
    for (int k = 0; k < 2; ++k) {
        if (ap.n[k] > 0) {
            const int rc = enqueue_finalize(c, k, ap.tasks[k], ap.n[k], ap.integrals[k], ap.results[k], s);
            if (rc) return rc;
            launch_row_scatter(ap.tasks[k], ap.results[k], k == 0 ? offS : offA, ap.rowLo, rows, weights, out, s);
        }
        if (level < 0 && refinements) launch_take_rows(ap.refCells + (size_t)k * c->nc, ap.rowLo, rows, refinements + (size_t)k * rows, s);
    }
    I2_CUDA(cudaGetLastError());
    if (stats) {
        QueueState q[2];
        unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 2; ++k) {
            std::memset(&q[k], 0, sizeof(QueueState));
            if (ap.n[k] > 0) I2_CUDA(cudaMemcpyAsync(&q[k], c->scr[k].qs, sizeof(QueueState), cudaMemcpyDeviceToHost, s));
        }
        if (level < 0) I2_CUDA(cudaMemcpyAsync(h, c->partial, sizeof(h), cudaMemcpyDeviceToHost, s));
        I2_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < 2; ++k)
            if (ap.n[k] > 0) fill_stats(q[k], ap.n[k], level, &stats[k]);
        if (level < 0) {
            int L;
            std::memcpy(&L, &h[6], sizeof(int));
            stats[2].last_round = L;
            stats[2].integrated[0] = (long long)h[0];
            long long before = (long long)h[0];
            for (int m = 1; m <= L && m <= MAX_REFINE_LEVEL; ++m) {
                stats[2].integrated[m] = before << (2 * m);
                stats[2].unconverged[m] = (long long)h[m];
                before = (long long)h[m];
            }
        }
    }
    return 0;
}

int i2_apply(i2_context *c, int level, const double *weights, double *out, unsigned char *refinements, i2_stats stats[3]) {
    int rc = i2_apply_rounds(c, level, weights);
    if (!rc) rc = i2_apply_finish(c, level, weights, out, refinements, stats);
    return rc;
}

int i2_symmetry_error(i2_context *c, const double *results, long long nHalf, double *errors) {
    if (!c || nHalf < 0 || (nHalf > 0 && (!results || !errors))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    launch_symmetry_error(results, nHalf, errors, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

int i2_error_summary(i2_context *c, const double *errors, long long n, double out[2]) {
    if (!c || !out || n < 0 || (n > 0 && !errors)) return I2_E_BADARG;
    out[0] = out[1] = 0.0;
    if (n == 0) return 0;
    I2_CUDA(cudaSetDevice(c->device));
    double *d = c->sumScratch + 12;
    I2_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 2, c->stream));
    launch_error_summary(errors, n, d, c->numSMs, c->stream);
    I2_CUDA(cudaMemcpyAsync(out, d, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    out[1] /= (double)n;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------------------
namespace {
inline long long align32(long long x) { return x & ~31LL; }
inline double *results_of(i2_context *c, int k) { return c->hResultsTarget[k] ? c->hResultsTarget[k] : c->hResults[k]; }
}  // namespace

// phase A of i2_host_prepare: upload, geometry, SoA pack, classification by vertex incidence -> pairs per class
extern "C++" int i2::host_prepare_mesh(i2_context *c, const double *hv, int nv, const int *hc, int nc, long long pairs[3]) {
    if (!c || !hv || !hc || nv <= 0 || nc <= 0 || !pairs) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // whatever an earlier prepare left behind is stale from here on (a failure below must not leave old lists next to a new mesh)
    c->hPrepared = false;
    for (int k = 0; k < 3; ++k) c->hCount[k] = c->hN[k] = c->hHalf[k] = c->hLo[k] = 0;
    // device buffers are kept between calls and only grow (cudaMalloc/cudaFree of multi-GB buffers costs milliseconds)
    int rc = ensure(&c->hVerts, &c->capVerts, (size_t)3 * nv);
    if (!rc) rc = ensure(&c->hCells, &c->capCells, (size_t)3 * nc);
    if (!rc) rc = ensure(&c->hNormals, &c->capNormals, (size_t)3 * nc);
    if (!rc) rc = ensure(&c->hMeasures, &c->capMeasures, (size_t)nc);
    if (!rc) rc = ensure(&c->hRefAll, &c->capRefAll, (size_t)3 * nc);
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) c->hRefinements[k] = c->hRefAll + (size_t)k * nc;
    I2_CUDA(cudaMemcpyAsync(c->hVerts, hv, sizeof(double) * 3 * nv, cudaMemcpyHostToDevice, s));
    I2_CUDA(cudaMemcpyAsync(c->hCells, hc, sizeof(int) * 3 * nc, cudaMemcpyHostToDevice, s));
    rc = i2_mesh_geometry(c, c->hVerts, nv, c->hCells, nc, c->hNormals, nullptr, c->hMeasures);
    if (rc) return rc;
    rc = i2_set_mesh(c, c->hVerts, nv, c->hCells, nc, c->hNormals, c->hMeasures);
    if (rc) return rc;
    // classification by vertex incidence: per-row partner counts -> first slot of every row in the three lists, totals
    rc = i2::ensure_incidence(c);
    if (rc) return rc;
    unsigned long long *totalsDev = c->rowOff + (size_t)3 * (nc + 1);
    unsigned long long totals[3];
    I2_CUDA(cudaMemcpyAsync(totals, totalsDev, sizeof(totals), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k) {
        pairs[k] = (long long)totals[k];
        c->hCount[k] = 2 * pairs[k];
    }
    c->hNv = nv;
    return 0;
}

// phase B: this context's shard of the three ordered task lists (needs host_prepare_mesh; ranges from i2_host_set_shard or
// host_set_forward_ranges)
extern "C++" int i2::host_prepare_lists(i2_context *c) {
    if (!c || !c->hVerts || c->nc <= 0) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int nc = c->nc, nv = c->hNv;
    int rc = 0;
    long long pairs[3];
    for (int k = 0; k < 3; ++k) {
        pairs[k] = c->hCount[k] / 2;
        if (2 * pairs[k] > INT_MAX) return I2_E_TOOBIG;
    }
    // this context's shard: the pairs with forward slots [lo, hi) of every class, in both orders; bounds are multiples of 32 so
    // that the warp groups of the shard are warp groups of the whole list (results do not depend on the partition)
    long long lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        if (c->explicitRanges) {
            lo[k] = align32(c->rangeLo[k] < pairs[k] ? c->rangeLo[k] : pairs[k]);
            hi[k] = c->rangeHi[k] >= pairs[k] ? pairs[k] : align32(c->rangeHi[k]);
        } else {
            lo[k] = align32(pairs[k] * c->shardRank / c->shardWorld);
            hi[k] = c->shardRank == c->shardWorld - 1 ? pairs[k] : align32(pairs[k] * (c->shardRank + 1) / c->shardWorld);
        }
        if (hi[k] < lo[k]) hi[k] = lo[k];
        const long long n = 2 * (hi[k] - lo[k]);
        rc = ensure(&c->hTasks[k], &c->capTasks[k], (size_t)3 * n);
        if (!rc) rc = ensure(&c->hIntegrals[k], &c->capIntegrals[k], (size_t)4 * n);
        if (!rc) rc = ensure(&c->hResults[k], &c->capResults[k], (size_t)3 * n);
        if (rc) return rc;
    }
    // adjacent classes: the whole ordered lists are small (~12 N and ~3 N tasks); a shard copies its two pieces out of them
    int *adj[2];
    bool whole[2];
    for (int k = 0; k < 2; ++k) {
        whole[k] = lo[k] == 0 && hi[k] == pairs[k];
        adj[k] = c->hTasks[k];
        if (!whole[k]) {
            rc = ensure(&c->adjFull[k], &c->capAdjFull[k], (size_t)6 * pairs[k]);
            if (rc) return rc;
            adj[k] = c->adjFull[k];
        }
    }
    launch_partners_fill(c->hCells, nv, nc, c->incScratch, c->rowOff, pairs[0] ? adj[0] : nullptr, pairs[1] ? adj[1] : nullptr, true, s);
    for (int k = 0; k < 2; ++k) {
        const long long h = hi[k] - lo[k];
        if (whole[k] || h == 0) continue;
        I2_CUDA(cudaMemcpyAsync(c->hTasks[k], adj[k] + 3 * lo[k], sizeof(int) * 3 * h, cudaMemcpyDeviceToDevice, s));
        I2_CUDA(cudaMemcpyAsync(c->hTasks[k] + 3 * h, adj[k] + 3 * (pairs[k] + lo[k]), sizeof(int) * 3 * h, cudaMemcpyDeviceToDevice, s));
    }
    // regular class: only this shard's slots are materialised, pairs and reversed pairs by one kernel
    launch_regular_fill(c->hCells, nc, c->rowOff, (unsigned long long)lo[2], (unsigned long long)hi[2], c->hTasks[2],
                        c->hTasks[2] + 3 * (hi[2] - lo[2]), c->numSMs, s);
    I2_CUDA(cudaGetLastError());
    // no synchronisation here: everything that consumes the lists is enqueued on the same stream
    for (int k = 0; k < 3; ++k) {
        c->hLo[k] = lo[k];
        c->hHalf[k] = hi[k] - lo[k];
        c->hN[k] = 2 * c->hHalf[k];
    }
    c->hPrepared = true;
    return 0;
}

int i2_host_prepare(i2_context *c, const double *hv, int nv, const int *hc, int nc, long long taskCounts[3]) {
    if (!taskCounts) return I2_E_BADARG;
    long long pairs[3];
    int rc = i2::host_prepare_mesh(c, hv, nv, hc, nc, pairs);
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) taskCounts[k] = 2 * pairs[k];
    return i2::host_prepare_lists(c);
}

int i2_host_set_shard(i2_context *c, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world) return I2_E_BADARG;
    c->shardRank = rank;
    c->shardWorld = world;
    c->explicitRanges = false;
    return 0;
}

extern "C++" int i2::host_set_forward_ranges(i2_context *c, const long long lo[3], const long long hi[3]) {
    if (!c || !lo || !hi) return I2_E_BADARG;
    for (int k = 0; k < 3; ++k) {
        if (lo[k] < 0 || hi[k] < lo[k]) return I2_E_BADARG;
        c->rangeLo[k] = lo[k];
        c->rangeHi[k] = hi[k];
    }
    c->explicitRanges = true;
    return 0;
}

int i2_host_shard(i2_context *c, long long first[3], long long count[3]) {
    if (!c || !first || !count) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    for (int k = 0; k < 3; ++k) {
        first[k] = c->hLo[k];
        count[k] = c->hN[k];
    }
    return 0;
}

int i2_host_row_costs(i2_context *c, int upper_only, double *h_cost, unsigned long long *h_row_first_regular) {
    if (!c) return I2_E_BADARG;
    if (!c->hVerts || !c->rowOff || !c->tri || c->nc <= 0) return I2_E_NOMESH;   // needs the mesh phase of the prepare only
    I2_CUDA(cudaSetDevice(c->device));
    const int nc = c->nc;
    if (h_cost) {
        int rc = ensure(&c->rowCost, &c->rowCostCap, (size_t)nc);
        if (rc) return rc;
        launch_row_cost(packed(c), upper_only != 0, c->rowCost, c->stream);
        I2_CUDA(cudaGetLastError());
        I2_CUDA(cudaMemcpyAsync(h_cost, c->rowCost, sizeof(double) * nc, cudaMemcpyDeviceToHost, c->stream));
    }
    if (h_row_first_regular)
        I2_CUDA(cudaMemcpyAsync(h_row_first_regular, c->rowOff + 2 * (size_t)(nc + 1), sizeof(unsigned long long) * (nc + 1), cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int i2_host_checksums(i2_context *c, double sums[12]) {
    if (!c || !sums) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    double *d = c->sumScratch;
    I2_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 12, c->stream));
    for (int k = 0; k < 3; ++k) launch_checksum(results_of(c, k), c->hN[k], d + 4 * k, c->numSMs, c->stream);
    I2_CUDA(cudaMemcpyAsync(sums, d, sizeof(double) * 12, cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int i2_host_device_views(i2_context *c, const int *tasks[3], const double *results[3]) {
    if (!c) return I2_E_BADARG;
    for (int k = 0; k < 3; ++k) {
        if (tasks) tasks[k] = c->hTasks[k];
        if (results) results[k] = results_of(c, k);
    }
    return 0;
}

// Integration rounds of the three classes of this context's shard: the two adjacent classes (short, latency-bound chains of
// kernels) on the side streams, the regular class on the context's stream, joined at the end.  At a fixed level the classes are
// complete afterwards (finalize on the same streams); under error control the finalize step is separate (host_run_finalize),
// so that a multi-GPU caller can agree on the last round first.
namespace {
int host_run_rounds_enqueue(i2_context *c, int level);
const bool g_useGraphs = [] { const char *e = getenv("I2_GRAPHS"); return !(e && atoi(e) == 0); }();
}

extern "C++" int i2::host_run_rounds(i2_context *c, int level) {
    if (!c) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (level > 12) return I2_E_LEVEL;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // Error control = a fixed chain of ~70 small dependent launches (23 per class) whose task counts live on the device: capture it
    // once per prepared shard and replay it (env I2_GRAPHS=0 turns this off).  Not on the legacy default stream (capture is not
    // allowed there: the drop-in classes) and not while profiling events are wanted.
    if (level < 0 && g_useGraphs && s != nullptr && !c->profiling) {
        int rc = i2_host_reserve(c, level, 0);      // every allocation the chain needs, before the capture (and before the key)
        if (rc) return rc;
        // a captured chain stays valid as long as every pointer and count it was recorded with is unchanged (a re-prepare of the same
        // mesh reuses the buffers): FNV-1a over all of them
        unsigned long long key = 1469598103934665603ull;
        auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
        mix((unsigned long long)(level + 7)); mix((unsigned long long)c->mathMode); mix(c->ruleShape13 ? 1 : 0); mix((unsigned long long)c->nc);
        mix((unsigned long long)c->stride); mix((uintptr_t)c->tri); mix((uintptr_t)c->cells); mix((uintptr_t)c->verts); mix((uintptr_t)c->hRefAll);
        for (int k = 0; k < 3; ++k) {
            const i2_context::ClassScratch &sc = c->scr[k];
            mix((uintptr_t)c->hTasks[k]); mix((uintptr_t)c->hIntegrals[k]); mix((uintptr_t)results_of(c, k)); mix((unsigned long long)c->hN[k]);
            mix((unsigned long long)c->hHalf[k]); mix((uintptr_t)sc.bufB); mix((uintptr_t)sc.rest[0]); mix((uintptr_t)sc.rest[1]);
            mix((uintptr_t)sc.cellFlag); mix((uintptr_t)sc.blockCnt); mix((uintptr_t)sc.qs);
        }
        if (!c->roundsGraph || c->roundsGraphKey != key) {
            if (c->roundsGraph) { cudaGraphExecDestroy(c->roundsGraph); c->roundsGraph = nullptr; }
            const long long before = g_launchCount;
            I2_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            rc = host_run_rounds_enqueue(c, level);
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(s, &graph);
            c->roundsGraphLaunches = g_launchCount - before;
            g_launchCount -= c->roundsGraphLaunches;     // captured, not launched: the replays count
            if (rc || e != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                return rc ? rc : (int)e;
            }
            const cudaError_t ei = cudaGraphInstantiate(&c->roundsGraph, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) { c->roundsGraph = nullptr; return (int)ei; }
            c->roundsGraphKey = key;
        }
        I2_CUDA(cudaGraphLaunch(c->roundsGraph, s));
        g_launchCount += c->roundsGraphLaunches;
        return 0;
    }
    return host_run_rounds_enqueue(c, level);
}

namespace {
int host_run_rounds_enqueue(i2_context *c, int level) {
    cudaStream_t s = c->stream;
    if (level < 0) I2_CUDA(cudaMemsetAsync(c->hRefAll, 0, (size_t)3 * c->nc, s));
    I2_CUDA(cudaEventRecord(c->forkEv, s));
    for (int k = 0; k < 3; ++k) {
        cudaStream_t st = k < 2 ? c->side[k] : s;
        if (k < 2) I2_CUDA(cudaStreamWaitEvent(st, c->forkEv, 0));
        if (c->hN[k] > 0) {
            bool fused = false;
            const bool prof = k == 2 && c->profiling && level >= 0;   // i2_set_profiling: device time of the regular class's kernels
            int rc = enqueue_rounds(c, k, c->hTasks[k], c->hN[k], c->hHalf[k], level, c->hIntegrals[k], results_of(c, k),
                                    level < 0 ? c->hRefinements[k] : nullptr, nullptr, st, prof, &fused,
                                    c->hResultsTarget[k] ? 1 : 0 /* results leave the GPU: full-sector stores */);
            if (!rc && level >= 0 && !fused) rc = enqueue_finalize(c, k, c->hTasks[k], c->hN[k], c->hIntegrals[k], results_of(c, k), st);
            if (rc) return rc;
            if (prof) I2_CUDA(cudaEventRecord(c->prof[2], st));
        } else {
            // an empty shard of this class (more GPUs than 32-task groups): its last round must still read 0, a multi-GPU
            // caller takes the maximum over the shards next
            I2_CUDA(cudaMemsetAsync(c->scr[k].qs, 0, sizeof(QueueState), st));
        }
        if (k < 2) I2_CUDA(cudaEventRecord(c->sideDone[k], st));
    }
    for (int k = 0; k < 2; ++k) I2_CUDA(cudaStreamWaitEvent(s, c->sideDone[k], 0));   // join
    return 0;
}
}  // namespace

extern "C++" int i2::host_run_finalize(i2_context *c, int level, bool wantErrors) {
    if (!c) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    for (int k = 0; k < 3; ++k) {
        if (c->hN[k] == 0) continue;
        if (level < 0) {
            const int rc = enqueue_finalize(c, k, c->hTasks[k], c->hN[k], c->hIntegrals[k], results_of(c, k), s);
            if (rc) return rc;
        }
        if (wantErrors) {
            // the shard holds every pair in both orders (slot t and hHalf + t): the (i,j)/(j,i) defect is local
            const int rc = ensure(&c->hErrors[k], &c->capErrors[k], (size_t)c->hN[k]);
            if (rc) return rc;
            launch_symmetry_error(results_of(c, k), c->hHalf[k], c->hErrors[k], s);
            I2_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

// Allocates what the next i2_host_run / i2_mgpu_run of this level will need (work-queue scratch, second result buffer, defect
// arrays), so that a caller who times the run — the drop-in classes print "Time for ... integration" like the reference, whose
// buffers are allocated before its timers start (src/evaluators/evaluator3d.cu:122-154) — does not time cudaMalloc.
int i2_host_reserve(i2_context *c, int level, int check) {
    if (!c) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    for (int k = 0; k < 3; ++k) {
        const long long n = c->hN[k];
        if (n == 0) continue;
        int rc = 0;
        if (check) rc = ensure(&c->hErrors[k], &c->capErrors[k], (size_t)n);
        if (!rc && level < 0) {
            i2_context::ClassScratch &sc = c->scr[k];
            rc = ensure(&sc.bufB, &sc.bufBCap, (size_t)4 * n);
            if (!rc && (size_t)n > sc.restCap) {
                size_t cap0 = sc.restCap, cap1 = sc.restCap;
                rc = ensure(&sc.rest[0], &cap0, (size_t)n);
                if (!rc) rc = ensure(&sc.rest[1], &cap1, (size_t)n);
                if (!rc) sc.restCap = (size_t)n;
            }
            if (!rc) rc = ensure(&sc.cellFlag, &sc.cellFlagCap, (size_t)c->nc);
            if (!rc) rc = ensure(&sc.blockCnt, &sc.blockCntCap, (size_t)kCompareMaxBlocks);
        }
        if (rc) return rc;
    }
    return 0;
}

// "bring your own communicator": the two halves of i2_host_run with access to what has to be agreed between the shards
int i2_host_run_rounds(i2_context *c, int level) { return i2::host_run_rounds(c, level); }
int i2_host_run_finalize(i2_context *c, int level, int check) { return i2::host_run_finalize(c, level, check != 0); }

int i2_host_last_rounds(i2_context *c, int last[3], int set) {
    if (!c || !last) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    for (int k = 0; k < 3; ++k) {
        if (set) I2_CUDA(cudaMemcpyAsync(&c->scr[k].qs->lastRound, &last[k], sizeof(int), cudaMemcpyHostToDevice, c->stream));
        else I2_CUDA(cudaMemcpyAsync(&last[k], &c->scr[k].qs->lastRound, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int i2_host_refinements(i2_context *c, unsigned char *ref, int set) {
    if (!c || !ref) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    if (set) I2_CUDA(cudaMemcpyAsync(c->hRefAll, ref, (size_t)3 * c->nc, cudaMemcpyHostToDevice, c->stream));
    else I2_CUDA(cudaMemcpyAsync(ref, c->hRefAll, (size_t)3 * c->nc, cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int i2_host_fetch(i2_context *c, int cls, int *hTasks, double *hResults, double *hErrors) {
    if (!c || cls < 0 || cls > 2) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    const long long n = c->hN[cls];
    if (n > 0) {
        if (hTasks) I2_CUDA(cudaMemcpyAsync(hTasks, c->hTasks[cls], sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
        if (hResults) I2_CUDA(cudaMemcpyAsync(hResults, results_of(c, cls), sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
        if (hErrors) {
            if (!c->hErrors[cls]) return I2_E_BADARG;   // the last run did not compute the defects
            I2_CUDA(cudaMemcpyAsync(hErrors, c->hErrors[cls], sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int i2_host_run(i2_context *c, int level, int *const hTasks[3], double *const hResults[3], double *const hErrors[3],
                unsigned char *const hRefinements[3], i2_stats hStats[3]) {
    if (!c) return I2_E_BADARG;
    if (!c->hPrepared) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (level > 12) return I2_E_LEVEL;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream, cs = c->copyStream;
    if (hStats) std::memset(hStats, 0, 3 * sizeof(i2_stats));
    for (int k = 0; k < 3; ++k) c->hResultsTarget[k] = nullptr;   // the single-context entry point always uses its own buffers
    bool wantErrors = false;
    for (int k = 0; k < 3; ++k) wantErrors = wantErrors || (hErrors && hErrors[k]);
    const long long n2 = c->hN[2];
    const bool copyOut2 = (hResults && hResults[2]) || (hTasks && hTasks[2]);

    auto copies = [&](int k, cudaStream_t st) -> int {
        const long long n = c->hN[k];
        if (n == 0) return 0;
        if (hErrors && hErrors[k]) I2_CUDA(cudaMemcpyAsync(hErrors[k], c->hErrors[k], sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (hResults && hResults[k]) I2_CUDA(cudaMemcpyAsync(hResults[k], c->hResults[k], sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
        if (hTasks && hTasks[k]) I2_CUDA(cudaMemcpyAsync(hTasks[k], c->hTasks[k], sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, st));
        return 0;
    };

    if (level >= 0 && !wantErrors && copyOut2 && n2 > 0) {
        // fixed level, per-pair results wanted on the host: tasks are independent -> the regular class is integrated chunk by
        // chunk and every finished chunk travels on the copy stream while the next one computes.  Chunks are cut inside each of
        // the two segments at multiples of 32 tasks, so they reproduce the warp groups (and the bits) of the unchunked run.
        I2_CUDA(cudaEventRecord(c->forkEv, s));
        for (int k = 0; k < 2; ++k) {
            cudaStream_t st = c->side[k];
            I2_CUDA(cudaStreamWaitEvent(st, c->forkEv, 0));
            if (c->hN[k] > 0) {
                int rc = enqueue_class(c, k, c->hTasks[k], c->hN[k], level, c->hIntegrals[k], c->hResults[k], nullptr, nullptr, st, false, c->hHalf[k]);
                if (!rc) rc = copies(k, st);
                if (rc) return rc;
            }
            I2_CUDA(cudaEventRecord(c->sideDone[k], st));
        }
        const long long chunkTasks = 1LL << 24;  // 16 Mi tasks: 384 MiB of Point3 per chunk
        int turn = 0;
        for (int seg = 0; seg < 2; ++seg) {
            const long long segLo = seg ? c->hHalf[2] : 0, segHi = seg ? n2 : c->hHalf[2];
            for (long long off = segLo; off < segHi; off += chunkTasks) {
                const long long m = segHi - off < chunkTasks ? segHi - off : chunkTasks;
                int rc = enqueue_class(c, 2, c->hTasks[2] + 3 * off, m, level, c->hIntegrals[2] + 4 * off, c->hResults[2] + 3 * off, nullptr,
                                       nullptr, s, false);
                if (rc) return rc;
                I2_CUDA(cudaEventRecord(c->chunkDone[turn], s));
                I2_CUDA(cudaStreamWaitEvent(cs, c->chunkDone[turn], 0));
                turn ^= 1;
                if (hResults && hResults[2])
                    I2_CUDA(cudaMemcpyAsync(hResults[2] + 3 * off, c->hResults[2] + 3 * off, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, cs));
                if (hTasks && hTasks[2])
                    I2_CUDA(cudaMemcpyAsync(hTasks[2] + 3 * off, c->hTasks[2] + 3 * off, sizeof(int) * 3 * m, cudaMemcpyDeviceToHost, cs));
            }
        }
        for (int k = 0; k < 2; ++k) I2_CUDA(cudaStreamWaitEvent(s, c->sideDone[k], 0));   // join
    } else {
        int rc = i2::host_run_rounds(c, level);
        if (!rc) rc = i2::host_run_finalize(c, level, wantErrors);
        for (int k = 0; k < 3 && !rc; ++k) rc = copies(k, s);
        if (rc) return rc;
        if (level < 0 && hRefinements)
            for (int k = 0; k < 3; ++k)
                if (hRefinements[k]) I2_CUDA(cudaMemcpyAsync(hRefinements[k], c->hRefinements[k], c->nc, cudaMemcpyDeviceToHost, s));
    }
    QueueState h[3];
    if (hStats)
        for (int k = 0; k < 3; ++k)
            if (c->hN[k] > 0) I2_CUDA(cudaMemcpyAsync(&h[k], c->scr[k].qs, sizeof(QueueState), cudaMemcpyDeviceToHost, s));
    I2_CUDA(cudaStreamSynchronize(s));
    I2_CUDA(cudaStreamSynchronize(cs));
    if (hStats)
        for (int k = 0; k < 3; ++k)
            if (c->hN[k] > 0) fill_stats(h[k], c->hN[k], level, &hStats[k]);
    return 0;
}

int i2_selftest_math(i2_context *c, int op, const double *a, const double *b, long long n, double *out) {
    if (!c || op < 0 || op > 3 || n < 0 || (n > 0 && (!a || !out || (op >= 2 && !b)))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    launch_selftest_math(op, a, b, n, out, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

// ---- multi-GPU export over NVLink peer stores: the owner's result array is mapped into the other processes, and the
//      kernels' final-assembly stores write into it directly (see include/i2_abi.h) ------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == I2_PEER_HANDLE_BYTES, "CUDA IPC handle size");

int i2_peer_alloc(i2_context *c, unsigned long long bytes, void **ptr, unsigned char handle[I2_PEER_HANDLE_BYTES]) {
    if (!c || !ptr || !handle || bytes == 0) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    void *p = nullptr;
    I2_CUDA(cudaMalloc(&p, (size_t)bytes));     // a whole allocation of its own: an IPC handle always maps from the base
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    std::memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return 0;
}

int i2_peer_open(i2_context *c, const unsigned char handle[I2_PEER_HANDLE_BYTES], void **ptr) {
    if (!c || !ptr || !handle) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    // cudaIpcMemLazyEnablePeerAccess: the mapping enables peer access between the two devices (NVLink / NVSwitch) on demand
    I2_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int i2_peer_close(i2_context *c, void *ptr) {
    if (!c || !ptr) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    I2_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

int i2_peer_free(i2_context *c, void *ptr) {
    if (!c || !ptr) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    I2_CUDA(cudaFree(ptr));
    return 0;
}

int i2_launch_count(long long *count) {
    if (!count) return I2_E_BADARG;
    *count = g_launchCount;
    return 0;
}

int i2_set_profiling(i2_context *c, int enabled) {
    if (!c) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    c->profiling = enabled != 0;
    if (c->profiling && !c->prof[0])
        for (int k = 0; k < 3; ++k) I2_CUDA(cudaEventCreate(&c->prof[k]));
    return 0;
}

int i2_profile_last(i2_context *c, float *msIntegrate, float *msFinalize) {
    if (!c || !c->prof[0]) return I2_E_BADARG;
    I2_CUDA(cudaEventSynchronize(c->prof[2]));
    if (msIntegrate) I2_CUDA(cudaEventElapsedTime(msIntegrate, c->prof[0], c->prof[1]));
    if (msFinalize) I2_CUDA(cudaEventElapsedTime(msFinalize, c->prof[1], c->prof[2]));
    return 0;
}

int i2_peak_dfma_three_operand(i2_context *c, double *tflops) {
    if (!c || !tflops) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    double *buf = nullptr;
    I2_CUDA(cudaMalloc((void **)&buf, sizeof(double) * 17));
    double seed[16];
    for (int k = 0; k < 16; ++k) seed[k] = (k < 8 ? 1.0000001 : 0.9999999) + 1e-9 * k;   // |b c| ~ 1: no overflow over the loop
    I2_CUDA(cudaMemcpyAsync(buf + 1, seed, sizeof(seed), cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0, e1;
    I2_CUDA(cudaEventCreate(&e0));
    I2_CUDA(cudaEventCreate(&e1));
    const int blocks = c->numSMs * 8, iters = 10000;
    float ms = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        I2_CUDA(cudaEventRecord(e0, c->stream));
        launch_peak_dfma3(buf, buf + 1, iters, blocks, c->stream);
        I2_CUDA(cudaEventRecord(e1, c->stream));
        I2_CUDA(cudaEventSynchronize(e1));
        I2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    *tflops = (double)blocks * 256.0 * iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return 0;
}

int i2_peak_dfma_with_integer(i2_context *c, int intPerDfma, double *tflops) {
    if (!c || !tflops || intPerDfma < 0 || intPerDfma > 3) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    double *sink = nullptr;
    I2_CUDA(cudaMalloc((void **)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    I2_CUDA(cudaEventCreate(&e0));
    I2_CUDA(cudaEventCreate(&e1));
    const int blocks = c->numSMs * 8, iters = 20000;
    float ms = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        I2_CUDA(cudaEventRecord(e0, c->stream));
        launch_peak_mix(intPerDfma, sink, iters, blocks, c->stream);
        I2_CUDA(cudaEventRecord(e1, c->stream));
        I2_CUDA(cudaEventSynchronize(e1));
        I2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    *tflops = (double)blocks * 256.0 * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return 0;
}

int i2_peak_rates(i2_context *c, double *dfmaTflops, double *mufuGops) {
    if (!c) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    double *sink = nullptr;
    I2_CUDA(cudaMalloc((void **)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    I2_CUDA(cudaEventCreate(&e0));
    I2_CUDA(cudaEventCreate(&e1));
    const int blocks = c->numSMs * 8, iters = 20000;
    float ms = 0.f;
    for (int rep = 0; rep < 3; ++rep) {  // last repetition is reported (first ones warm up clocks)
        I2_CUDA(cudaEventRecord(e0, c->stream));
        launch_peak_dfma(sink, iters, blocks, c->stream);
        I2_CUDA(cudaEventRecord(e1, c->stream));
        I2_CUDA(cudaEventSynchronize(e1));
        I2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    if (dfmaTflops) *dfmaTflops = (double)blocks * 256.0 * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
    for (int rep = 0; rep < 3; ++rep) {
        I2_CUDA(cudaEventRecord(e0, c->stream));
        launch_peak_mufu(sink, iters / 4, blocks, c->stream);
        I2_CUDA(cudaEventRecord(e1, c->stream));
        I2_CUDA(cudaEventSynchronize(e1));
        I2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    if (mufuGops) *mufuGops = (double)blocks * 256.0 * (iters / 4) * 4.0 / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return 0;
}

}  // extern "C"
