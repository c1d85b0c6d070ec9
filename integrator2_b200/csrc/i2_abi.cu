// C-ABI entry points of libintegrator2_b200.so (declared in include/i2_abi.h).
// Host-side orchestration only: every numerical step is a kernel in i2_kernels.cu; there is no CPU fallback.
#include "../../include/i2_abi.h"
#include "i2_kernels.cuh"

#include <climits>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace i2;

#define I2_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return (int)e__;        \
    } while (0)

struct i2_context {
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;
    bool ownStream = false;
    int mathMode = I2_MATH_FAST;
    bool haveQuad = false;

    // borrowed mesh arrays + owned SoA pack
    const double *verts = nullptr;
    const int *cells = nullptr;
    int nv = 0, nc = 0;
    double *tri = nullptr;
    int stride = 0;
    size_t triCap = 0;

    // work queue scratch, one set per neighbour class so that the three classes can run concurrently (owned, grown on demand)
    struct ClassScratch {
        double *bufB = nullptr;
        size_t bufBCap = 0;
        int *rest[2] = {nullptr, nullptr};   // [0] = dense list of unconverged slots (input order), [1] = per-CTA staging segments
        size_t restCap = 0;
        int *blockCnt = nullptr;             // per-CTA counts of the deterministic compaction
        size_t blockCntCap = 0;
        unsigned char *cellFlag = nullptr;
        size_t cellFlagCap = 0;
        QueueState *qs = nullptr;
    } scr[3];
    // i2_integrate_all / i2_host_run: the two adjacent classes run on side streams next to the regular class
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t forkEv = nullptr, sideDone[2] = {nullptr, nullptr};

    // matrix-free path scratch (per-chunk partial row sums)
    double *partial = nullptr;
    size_t partialCap = 0;
    unsigned char *depthBuf = nullptr;   // list-free adaptive path: per (column chunk, row) refinement depth
    size_t depthCap = 0;

    // classification scratch
    unsigned long long *rowCounts = nullptr;
    size_t rowCap = 0;

    // host-entry state (i2_host_prepare / i2_host_run)
    double *hVerts = nullptr, *hNormals = nullptr, *hMeasures = nullptr;
    int *hCells = nullptr;
    int *hTasks[3] = {nullptr, nullptr, nullptr};
    double *hIntegrals[3] = {nullptr, nullptr, nullptr};
    double *hResults[3] = {nullptr, nullptr, nullptr};
    double *hErrors[3] = {nullptr, nullptr, nullptr};
    unsigned char *hRefinements[3] = {nullptr, nullptr, nullptr};
    long long hCount[3] = {0, 0, 0};
    // multi-GPU use of the host-entry path: this context integrates slots [hLo, hLo + hN) of every class
    int shardRank = 0, shardWorld = 1;
    long long hLo[3] = {0, 0, 0}, hN[3] = {0, 0, 0};
    size_t capVerts = 0, capCells = 0, capNormals = 0, capMeasures = 0, capTasks[3] = {0, 0, 0}, capIntegrals[3] = {0, 0, 0},
           capResults[3] = {0, 0, 0}, capRefinements[3] = {0, 0, 0}, capErrors[3] = {0, 0, 0};
    cudaEvent_t chunkDone[2] = {nullptr, nullptr};
    bool profiling = false;
    cudaEvent_t prof[3] = {nullptr, nullptr, nullptr};
};

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
        fr(c->hErrors[k], c->capErrors[k]); fr(c->hRefinements[k], c->capRefinements[k]);
        c->hCount[k] = 0;
    }
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
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaStreamCreateWithFlags(&c->side[k], cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->sideDone[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->forkEv, cudaEventDisableTiming);
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
    for (auto &sc : c->scr) {
        if (sc.bufB) cudaFree(sc.bufB);
        for (int k = 0; k < 2; ++k) if (sc.rest[k]) cudaFree(sc.rest[k]);
        if (sc.cellFlag) cudaFree(sc.cellFlag);
        if (sc.blockCnt) cudaFree(sc.blockCnt);
        if (sc.qs) cudaFree(sc.qs);
    }
    for (int k = 0; k < 3; ++k) if (c->prof[k]) cudaEventDestroy(c->prof[k]);
    if (c->rowCounts) cudaFree(c->rowCounts);
    if (c->partial) cudaFree(c->partial);
    if (c->depthBuf) cudaFree(c->depthBuf);
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
    I2_CUDA(upload_quadrature(packed4, n, (double)(1 << order), c->stream));
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

int i2_classify_count(i2_context *c, const int *cells, int nc, long long counts[3]) {
    if (!c || !counts || nc < 0 || (nc > 0 && !cells)) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    counts[0] = counts[1] = counts[2] = 0;
    if (nc == 0) return 0;
    int rc = ensure(&c->rowCounts, &c->rowCap, (size_t)3 * nc);
    if (rc) return rc;
    launch_classify_count(cells, nc, c->rowCounts, c->stream);
    I2_CUDA(cudaGetLastError());
    std::vector<unsigned long long> h((size_t)3 * nc);
    I2_CUDA(cudaMemcpyAsync(h.data(), c->rowCounts, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long run[3] = {0, 0, 0};
    for (int i = 0; i < nc; ++i)
        for (int k = 0; k < 3; ++k) {
            const unsigned long long v = h[(size_t)3 * i + k];
            h[(size_t)3 * i + k] = run[k];  // exclusive prefix = first slot of row i
            run[k] += v;
        }
    I2_CUDA(cudaMemcpyAsync(c->rowCounts, h.data(), h.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; ++k) counts[k] = (long long)run[k];
    return 0;
}

int i2_classify_fill(i2_context *c, const int *cells, int nc, int *simple, int *attached, int *notn) {
    if (!c || nc < 0 || (nc > 0 && (!cells || !c->rowCounts))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    launch_classify_fill(cells, nc, c->rowCounts, simple, attached, notn, c->stream);
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

// enqueues the whole of one class on stream s (no synchronisation); uses the class's own scratch
int enqueue_class(i2_context *c, int cls, const int *tasks, long long n, int level, double *integrals, double *results,
                  unsigned char *refinements, unsigned char *converged, cudaStream_t s, bool profile) {
    i2_context::ClassScratch &sc = c->scr[cls];
    const PackedMesh pm = packed(c);
    I2_CUDA(cudaMemsetAsync(sc.qs, 0, sizeof(QueueState), s));
    bool fused = false;

    if (level >= 0) {
        if (profile) I2_CUDA(cudaEventRecord(c->prof[0], s));
        // regular pairs with the grouped kernel: the final assembly is fused into the integrate kernel
        fused = (cls == 2 && c->mathMode == I2_MATH_FAST);
        launch_integrate(cls, c->mathMode, pm, tasks, nullptr, nullptr, n, level, integrals, fused ? results : nullptr, c->numSMs, s);
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
        launch_integrate(cls, c->mathMode, pm, tasks, nullptr, nullptr, n, 0, integrals, nullptr, c->numSMs, s);
        launch_flag_cells(tasks, n, sc.cellFlag, s);
        launch_bump(sc.cellFlag, refinements, c->nc, s);
        // rounds 1..5 are enqueued unconditionally; a round whose device-side task count is 0 does nothing.
        for (int m = 1; m <= MAX_REFINE_LEVEL; ++m) {
            double *cur = (m & 1) ? sc.bufB : integrals;
            const double *prev = (m & 1) ? integrals : sc.bufB;
            const int *listIn = m == 1 ? nullptr : sc.rest[0];
            const int *countIn = m == 1 ? nullptr : &sc.qs->count[m - 1];
            launch_integrate(cls, c->mathMode, pm, tasks, listIn, countIn, n, m, cur, nullptr, c->numSMs, s);
            launch_compare(cur, prev, tasks, listIn, countIn, n, sc.rest[1], sc.blockCnt, sc.rest[0], &sc.qs->count[m], sc.cellFlag, converged, sc.qs, m,
                           c->numSMs, s);
            launch_bump(sc.cellFlag, refinements, c->nc, s);
        }
    }
    if (!fused) launch_finalize(cls, pm, c->verts, tasks, n, integrals, sc.bufB, sc.qs, results, sc.qs, s);
    I2_CUDA(cudaGetLastError());
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

int i2_integrate_class(i2_context *c, int cls, const int *tasks, long long n, int level, double *integrals, double *results,
                       unsigned char *refinements, unsigned char *converged, i2_stats *stats) {
    if (!c) return I2_E_BADARG;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    int rc = check_class_args(c, cls, tasks, n, level, integrals, results);
    if (rc || n == 0) return rc;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    rc = enqueue_class(c, cls, tasks, n, level, integrals, results, refinements, converged, s, c->profiling);
    if (rc) return rc;
    if (stats) {
        QueueState h;
        I2_CUDA(cudaMemcpyAsync(&h, c->scr[cls].qs, sizeof(h), cudaMemcpyDeviceToHost, s));
        I2_CUDA(cudaStreamSynchronize(s));
        fill_stats(h, n, level, stats);
    }
    return 0;
}

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
    // enough CTAs for a few waves of 4 CTAs/SM: split the columns when there are few row blocks
    const int rowBlocks = (rows + kThreads - 1) / kThreads;
    int chunks = (c->numSMs * 4 * 2 + rowBlocks - 1) / rowBlocks;
    if (chunks < 1) chunks = 1;
    const int maxChunks = (c->nc + 255) / 256;
    if (chunks > maxChunks) chunks = maxChunks;
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
    // 4 lanes per row -> 32 rows per CTA; enough CTAs for a few waves of 3 CTAs/SM: split the columns when there are few row blocks
    const int rowBlocks = (rows + kThreads / 4 - 1) / (kThreads / 4);
    int chunks = (c->numSMs * 3 * 2 + rowBlocks - 1) / rowBlocks;
    if (chunks < 1) chunks = 1;
    const int maxChunks = (c->nc + 255) / 256;
    if (chunks > maxChunks) chunks = maxChunks;
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

int i2_symmetry_error(i2_context *c, const double *results, long long nHalf, double *errors) {
    if (!c || nHalf < 0 || (nHalf > 0 && (!results || !errors))) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    launch_symmetry_error(results, nHalf, errors, c->stream);
    I2_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------------------
int i2_host_prepare(i2_context *c, const double *hv, int nv, const int *hc, int nc, long long taskCounts[3]) {
    if (!c || !hv || !hc || nv <= 0 || nc <= 0 || !taskCounts) return I2_E_BADARG;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // device buffers are kept between calls and only grow (cudaMalloc/cudaFree of multi-GB buffers costs milliseconds)
    int rc = ensure(&c->hVerts, &c->capVerts, (size_t)3 * nv);
    if (!rc) rc = ensure(&c->hCells, &c->capCells, (size_t)3 * nc);
    if (!rc) rc = ensure(&c->hNormals, &c->capNormals, (size_t)3 * nc);
    if (!rc) rc = ensure(&c->hMeasures, &c->capMeasures, (size_t)nc);
    if (rc) return rc;
    I2_CUDA(cudaMemcpyAsync(c->hVerts, hv, sizeof(double) * 3 * nv, cudaMemcpyHostToDevice, s));
    I2_CUDA(cudaMemcpyAsync(c->hCells, hc, sizeof(int) * 3 * nc, cudaMemcpyHostToDevice, s));
    rc = i2_mesh_geometry(c, c->hVerts, nv, c->hCells, nc, c->hNormals, nullptr, c->hMeasures);
    if (rc) return rc;
    rc = i2_set_mesh(c, c->hVerts, nv, c->hCells, nc, c->hNormals, c->hMeasures);
    if (rc) return rc;
    long long pairs[3];
    rc = i2_classify_count(c, c->hCells, nc, pairs);
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) {
        if (2 * pairs[k] > INT_MAX) return I2_E_TOOBIG;
        c->hCount[k] = 2 * pairs[k];
        taskCounts[k] = c->hCount[k];
        rc = ensure(&c->hTasks[k], &c->capTasks[k], (size_t)3 * c->hCount[k]);
        if (!rc) rc = ensure(&c->hIntegrals[k], &c->capIntegrals[k], (size_t)4 * c->hCount[k]);
        if (!rc) rc = ensure(&c->hResults[k], &c->capResults[k], (size_t)3 * c->hCount[k]);
        if (!rc) rc = ensure(&c->hRefinements[k], &c->capRefinements[k], (size_t)nc);
        if (rc) return rc;
    }
    rc = i2_classify_fill(c, c->hCells, nc, c->hTasks[0], c->hTasks[1], c->hTasks[2]);
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) {
        rc = i2_add_reversed_pairs(c, c->hTasks[k], pairs[k]);
        if (rc) return rc;
    }
    for (int k = 0; k < 3; ++k) {   // contiguous equal-count shard of this context (the whole list unless i2_host_set_shard was called)
        const long long base = c->hCount[k] / c->shardWorld, rem = c->hCount[k] % c->shardWorld;
        c->hLo[k] = base * c->shardRank + (c->shardRank < rem ? c->shardRank : rem);
        c->hN[k] = base + (c->shardRank < rem ? 1 : 0);
    }
    I2_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int i2_host_set_shard(i2_context *c, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world) return I2_E_BADARG;
    c->shardRank = rank;
    c->shardWorld = world;
    return 0;
}

int i2_host_shard(i2_context *c, long long first[3], long long count[3]) {
    if (!c || !first || !count) return I2_E_BADARG;
    if (!c->hVerts) return I2_E_NOMESH;
    for (int k = 0; k < 3; ++k) {
        first[k] = c->hLo[k];
        count[k] = c->hN[k];
    }
    return 0;
}

int i2_host_checksums(i2_context *c, double sums[12]) {
    if (!c || !sums) return I2_E_BADARG;
    if (!c->hVerts) return I2_E_NOMESH;
    I2_CUDA(cudaSetDevice(c->device));
    double *d = nullptr;
    I2_CUDA(cudaMalloc((void **)&d, sizeof(double) * 12));
    I2_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 12, c->stream));
    for (int k = 0; k < 3; ++k) launch_checksum(c->hResults[k] + 3 * c->hLo[k], c->hN[k], d + 4 * k, c->numSMs, c->stream);
    I2_CUDA(cudaMemcpyAsync(sums, d, sizeof(double) * 12, cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

int i2_host_device_views(i2_context *c, const int *tasks[3], const double *results[3]) {
    if (!c) return I2_E_BADARG;
    for (int k = 0; k < 3; ++k) {
        if (tasks) tasks[k] = c->hTasks[k];
        if (results) results[k] = c->hResults[k];
    }
    return 0;
}

int i2_host_run(i2_context *c, int level, int *const hTasks[3], double *const hResults[3], double *const hErrors[3],
                unsigned char *const hRefinements[3], i2_stats hStats[3]) {
    if (!c) return I2_E_BADARG;
    if (!c->hVerts) return I2_E_NOMESH;
    if (!c->haveQuad) return I2_E_NOQUAD;
    if (level > 12) return I2_E_LEVEL;
    I2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream, cs = c->copyStream;
    if (hStats) std::memset(hStats, 0, 3 * sizeof(i2_stats));

    // one whole class on stream st: integration, optional (i,j)/(j,i) defect, device-to-host copies; no synchronisation
    // device views of this context's shard of class k
    auto dTasks = [&](int k) { return c->hTasks[k] + 3 * c->hLo[k]; };
    auto dIntegrals = [&](int k) { return c->hIntegrals[k] + 4 * c->hLo[k]; };
    auto dResults = [&](int k) { return c->hResults[k] + 3 * c->hLo[k]; };
    auto whole = [&](int k, cudaStream_t st) -> int {
        const long long n = c->hN[k];
        int rc = enqueue_class(c, k, dTasks(k), n, level, dIntegrals(k), dResults(k), level < 0 ? c->hRefinements[k] : nullptr,
                               nullptr, st, false);
        if (rc) return rc;
        if (hErrors && hErrors[k]) {
            if (c->shardWorld > 1) return I2_E_BADARG;   // the (i,j)/(j,i) defect pairs slot t with slot n/2 + t: whole lists only
            rc = ensure(&c->hErrors[k], &c->capErrors[k], (size_t)n);
            if (rc) return rc;
            launch_symmetry_error(c->hResults[k], n / 2, c->hErrors[k], st);
            I2_CUDA(cudaGetLastError());
            I2_CUDA(cudaMemcpyAsync(hErrors[k], c->hErrors[k], sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        }
        if (hResults && hResults[k])
            I2_CUDA(cudaMemcpyAsync(hResults[k], dResults(k), sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
        if (hTasks && hTasks[k])
            I2_CUDA(cudaMemcpyAsync(hTasks[k], dTasks(k), sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, st));
        if (level < 0 && hRefinements && hRefinements[k])
            I2_CUDA(cudaMemcpyAsync(hRefinements[k], c->hRefinements[k], c->nc, cudaMemcpyDeviceToHost, st));
        return 0;
    };

    // fork: the adjacent classes run on the side streams, concurrently with the regular class on the context's stream
    I2_CUDA(cudaEventRecord(c->forkEv, s));
    for (int k = 0; k < 2; ++k) {
        cudaStream_t st = c->side[k];
        I2_CUDA(cudaStreamWaitEvent(st, c->forkEv, 0));
        if (level < 0) I2_CUDA(cudaMemsetAsync(c->hRefinements[k], 0, c->nc, st));
        if (c->hN[k] > 0) {
            const int rc = whole(k, st);
            if (rc) return rc;
        }
        I2_CUDA(cudaEventRecord(c->sideDone[k], st));
    }
    {
        const int k = 2;
        const long long n = c->hN[k];
        if (level < 0) I2_CUDA(cudaMemsetAsync(c->hRefinements[k], 0, c->nc, s));
        const bool wantErr = hErrors && hErrors[k];
        if (n > 0 && level >= 0 && !wantErr) {
            // fixed level: tasks are independent -> integrate chunk by chunk, copy each finished chunk on the copy stream
            const long long chunkTasks = 1LL << 24;  // 16 Mi tasks: 384 MiB of Point3 per chunk
            const bool copyOut = (hResults && hResults[k]) || (hTasks && hTasks[k]);
            int turn = 0;
            for (long long off = 0; off < n; off += copyOut ? chunkTasks : n) {
                const long long m = !copyOut ? n : ((n - off < chunkTasks) ? (n - off) : chunkTasks);
                int rc = enqueue_class(c, k, dTasks(k) + 3 * off, m, level, dIntegrals(k) + 4 * off, dResults(k) + 3 * off,
                                       nullptr, nullptr, s, false);
                if (rc) return rc;
                if (!copyOut) break;
                I2_CUDA(cudaEventRecord(c->chunkDone[turn], s));
                I2_CUDA(cudaStreamWaitEvent(cs, c->chunkDone[turn], 0));
                turn ^= 1;
                if (hResults && hResults[k])
                    I2_CUDA(cudaMemcpyAsync(hResults[k] + 3 * off, dResults(k) + 3 * off, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, cs));
                if (hTasks && hTasks[k])
                    I2_CUDA(cudaMemcpyAsync(hTasks[k] + 3 * off, dTasks(k) + 3 * off, sizeof(int) * 3 * m, cudaMemcpyDeviceToHost, cs));
            }
        } else if (n > 0) {
            const int rc = whole(k, s);
            if (rc) return rc;
        }
    }
    for (int k = 0; k < 2; ++k) I2_CUDA(cudaStreamWaitEvent(s, c->sideDone[k], 0));   // join
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
