// Multi-GPU layer of the C ABI (i2_mgpu_* in include/i2_abi.h): the pair lists of Evaluator3D::runAllPairs
// (src/evaluators/evaluator3d.cu:120-204) sharded over the GPUs of one box.  The reference is single-GPU; there is nothing
// to match but the results.
//
// One i2_mgpu handle drives the local GPUs of this process: all GPUs of the box from one process (i2_mgpu_create_local: the CLI
// and the drop-in host classes, env I2_GPUS) or one GPU per process (i2_mgpu_create_rank: torchrun-style launches).  Every GPU
// classifies the mesh by vertex incidence (O(N valence^2), replicated) and materialises ONLY ITS SHARD of the task lists: the
// pairs with forward slots [lo, hi) of each class, in both orders ([pairs ; reversed pairs], the shape of a small runAllPairs
// list), lo and hi multiples of 32 so that the warp groups — and therefore the bits of every result — are those of the
// unsharded run.  All descendants of a task stay on its GPU and a pair's (i,j)/(j,i) defect is local.
// Exchange steps (NCCL, loaded with dlopen: no link-time dependency, none at all for a single GPU):
//   error control : all-reduce(max) of each class's last round L before the final assembly — the reference's result ping-pong
//                    makes a converged pair's value depend on the parity of the GLOBAL L — and all-reduce(max) of
//                   the per-cell refinement counters (a cell's tasks may sit on several GPUs; the rounds a cell is flagged in
//                   form a prefix, so the maximum over the GPUs is the count of the unsharded run);
//   statistics    : all-reduce(sum) of the per-round counts and of the checksums;
//   export        : i2_mgpu_gather (send/recv of the result shards to one GPU) for callers that want everything in one place;
//                   the default is row-striped: every GPU keeps, formats and writes the rows it computed (i2_mgpu_fetch).
#include "i2_context.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace i2;

#define I2_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return (int)e__;        \
    } while (0)
#define I2_NCCL(call)                                                                                      \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess) {                                                                          \
            fprintf(stderr, "i2_mgpu: NCCL error %d (%s) at %s:%d\n", (int)r__, api->GetErrorString(r__), __FILE__, __LINE__); \
            return I2_E_NCCL;                                                                              \
        }                                                                                                  \
    } while (0)

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

// libnccl.so.2 by SONAME: inside a process that already loaded NCCL (e.g. through torch) this returns that copy
NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char *names[] = {getenv("I2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names)
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) {
        fprintf(stderr, "i2_mgpu: cannot load libnccl.so.2 (%s); set I2_NCCL_LIB\n", dlerror());
        return nullptr;
    }
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) { fprintf(stderr, "i2_mgpu: %s missing in NCCL\n", n); ok = false; } return p; };
    *(void **)&api.GetUniqueId = sym("ncclGetUniqueId");
    *(void **)&api.CommInitRank = sym("ncclCommInitRank");
    *(void **)&api.CommInitAll = sym("ncclCommInitAll");
    *(void **)&api.CommDestroy = sym("ncclCommDestroy");
    *(void **)&api.AllReduce = sym("ncclAllReduce");
    *(void **)&api.Send = sym("ncclSend");
    *(void **)&api.Recv = sym("ncclRecv");
    *(void **)&api.GroupStart = sym("ncclGroupStart");
    *(void **)&api.GroupEnd = sym("ncclGroupEnd");
    *(void **)&api.GetErrorString = sym("ncclGetErrorString");
    if (!ok) return nullptr;
    api.handle = h;
    return &api;
}

inline long long align32(long long x) { return x & ~31LL; }

}  // namespace

static_assert(sizeof(ncclUniqueId) == I2_MGPU_ID_BYTES, "NCCL unique id size");

struct i2_mgpu {
    int world = 1, nLocal = 1, firstRank = 0;
    std::vector<i2_context *> ctx;
    std::vector<ncclComm_t> comm;
    std::vector<double *> scratch;        // per local GPU: 64 doubles of device scratch for the small all-reduces
    std::vector<double *> target[3];      // per class, per local GPU: results target override (NULL: the context's own buffer)
    long long pairs[3] = {0, 0, 0};
    std::vector<long long> lo[3], hi[3];  // forward-slot range of every rank of the job, per class
    int nc = 0;
    bool prepared = false;
    int lastLevel = 0;
    // operator apply: row blocks of all ranks, per local GPU the full result vector, the weights and the block's counters
    std::vector<int> rowCut;
    std::vector<double *> apFull, apW;
    std::vector<unsigned char *> apRef;
    bool applyPrepared = false;
};

namespace {

// runs fn(k) for every local GPU, concurrently when there are several (prepare and the adaptive chains synchronise / launch
// dozens of kernels per GPU); returns the first error
template <class F>
int for_local(i2_mgpu *mg, F fn) {
    if (mg->nLocal == 1) return fn(0);
    std::vector<int> rc(mg->nLocal, 0);
    std::vector<std::thread> th;
    th.reserve(mg->nLocal);
    for (int k = 0; k < mg->nLocal; ++k) th.emplace_back([&, k] { rc[k] = fn(k); });
    for (auto &t : th) t.join();
    for (int r : rc)
        if (r) return r;
    return 0;
}

int destroy_partial(i2_mgpu *mg) {
    NcclApi *api = mg->world > 1 ? nccl_api() : nullptr;
    for (size_t k = 0; k < mg->ctx.size(); ++k) {
        if (mg->ctx[k]) cudaSetDevice(mg->ctx[k]->device);
        if (k < mg->comm.size() && mg->comm[k] && api) api->CommDestroy(mg->comm[k]);
        if (k < mg->scratch.size() && mg->scratch[k]) cudaFree(mg->scratch[k]);
        if (k < mg->apFull.size() && mg->apFull[k]) cudaFree(mg->apFull[k]);
        if (k < mg->apW.size() && mg->apW[k]) cudaFree(mg->apW[k]);
        if (k < mg->apRef.size() && mg->apRef[k]) cudaFree(mg->apRef[k]);
        if (mg->ctx[k]) i2_destroy(mg->ctx[k]);
    }
    delete mg;
    return 0;
}

int finish_create(i2_mgpu *mg) {
    mg->scratch.assign(mg->nLocal, nullptr);
    mg->apFull.assign(mg->nLocal, nullptr);
    mg->apW.assign(mg->nLocal, nullptr);
    mg->apRef.assign(mg->nLocal, nullptr);
    for (int c = 0; c < 3; ++c) mg->target[c].assign(mg->nLocal, nullptr);
    for (int k = 0; k < mg->nLocal; ++k) {
        I2_CUDA(cudaSetDevice(mg->ctx[k]->device));
        I2_CUDA(cudaMalloc((void **)&mg->scratch[k], sizeof(double) * 64));
    }
    for (int c = 0; c < 3; ++c) { mg->lo[c].assign(mg->world, 0); mg->hi[c].assign(mg->world, 0); }
    return 0;
}

}  // namespace

extern "C" {

int i2_mgpu_unique_id(unsigned char id[I2_MGPU_ID_BYTES]) {
    if (!id) return I2_E_BADARG;
    NcclApi *api = nccl_api();
    if (!api) return I2_E_NCCL;
    ncclUniqueId u;
    I2_NCCL(api->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return 0;
}

int i2_mgpu_create_rank(i2_mgpu **out, int device, int rank, int world, const unsigned char id[I2_MGPU_ID_BYTES]) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return I2_E_BADARG;
    *out = nullptr;
    i2_mgpu *mg = new i2_mgpu;
    mg->world = world; mg->nLocal = 1; mg->firstRank = rank;
    mg->ctx.assign(1, nullptr);
    mg->comm.assign(1, nullptr);
    int rc = i2_create(&mg->ctx[0], device);
    if (!rc && world > 1) {
        NcclApi *api = nccl_api();
        if (!api) rc = I2_E_NCCL;
        else {
            ncclUniqueId u;
            std::memcpy(&u, id, sizeof(u));
            cudaSetDevice(device);
            const ncclResult_t r = api->CommInitRank(&mg->comm[0], world, u, rank);
            if (r != ncclSuccess) { fprintf(stderr, "i2_mgpu: ncclCommInitRank: %s\n", api->GetErrorString(r)); rc = I2_E_NCCL; }
        }
    }
    if (!rc) rc = finish_create(mg);
    if (rc) { destroy_partial(mg); return rc; }
    *out = mg;
    return 0;
}

int i2_mgpu_create_local(i2_mgpu **out, int ngpus, const int *devices) {
    if (!out || ngpus < 1) return I2_E_BADARG;
    *out = nullptr;
    int have = 0;
    I2_CUDA(cudaGetDeviceCount(&have));
    std::vector<int> dev(ngpus);
    for (int k = 0; k < ngpus; ++k) {
        dev[k] = devices ? devices[k] : k;
        if (dev[k] < 0 || dev[k] >= have) return I2_E_BADARG;
    }
    i2_mgpu *mg = new i2_mgpu;
    mg->world = ngpus; mg->nLocal = ngpus; mg->firstRank = 0;
    mg->ctx.assign(ngpus, nullptr);
    mg->comm.assign(ngpus, nullptr);
    int rc = 0;
    for (int k = 0; k < ngpus && !rc; ++k) rc = i2_create(&mg->ctx[k], dev[k]);
    if (!rc && ngpus > 1) {
        NcclApi *api = nccl_api();
        if (!api) rc = I2_E_NCCL;
        else {
            const ncclResult_t r = api->CommInitAll(mg->comm.data(), ngpus, dev.data());
            if (r != ncclSuccess) { fprintf(stderr, "i2_mgpu: ncclCommInitAll: %s\n", api->GetErrorString(r)); rc = I2_E_NCCL; }
        }
    }
    if (!rc) rc = finish_create(mg);
    if (rc) { destroy_partial(mg); return rc; }
    *out = mg;
    return 0;
}

int i2_mgpu_destroy(i2_mgpu *mg) {
    if (!mg) return 0;
    for (int k = 0; k < mg->nLocal; ++k)
        if (mg->ctx[k]) i2_synchronize(mg->ctx[k]);
    return destroy_partial(mg);
}

int i2_mgpu_info(i2_mgpu *mg, int *world, int *n_local, int *first_rank) {
    if (!mg) return I2_E_BADARG;
    if (world) *world = mg->world;
    if (n_local) *n_local = mg->nLocal;
    if (first_rank) *first_rank = mg->firstRank;
    return 0;
}

i2_context *i2_mgpu_context(i2_mgpu *mg, int local_index) {
    if (!mg || local_index < 0 || local_index >= mg->nLocal) return nullptr;
    return mg->ctx[local_index];
}

int i2_mgpu_set_quadrature(i2_mgpu *mg, const double *xy, const double *w, int n, int order) {
    if (!mg) return I2_E_BADARG;
    for (int k = 0; k < mg->nLocal; ++k) {
        const int rc = i2_set_quadrature(mg->ctx[k], xy, w, n, order);
        if (rc) return rc;
    }
    return 0;
}

int i2_mgpu_set_math_mode(i2_mgpu *mg, int mode) {
    if (!mg) return I2_E_BADARG;
    for (int k = 0; k < mg->nLocal; ++k) {
        const int rc = i2_set_math_mode(mg->ctx[k], mode);
        if (rc) return rc;
    }
    return 0;
}

int i2_mgpu_synchronize(i2_mgpu *mg) {
    if (!mg) return I2_E_BADARG;
    for (int k = 0; k < mg->nLocal; ++k) {
        const int rc = i2_synchronize(mg->ctx[k]);
        if (rc) return rc;
    }
    return 0;
}

// Sharded prepare.  level < 0 (error control) balances the regular class by PREDICTED cost: rows are weighted with the
// expected number of child integrations of their pairs (k_row_cost) and the cuts are placed where the cumulative cost crosses
// r / world; at a fixed level every pair costs the same and the cuts are equal counts.  Every process computes the same cuts.
int i2_mgpu_prepare(i2_mgpu *mg, const double *hv, int nv, const int *hc, int nc, int level, long long taskCounts[3]) {
    if (!mg || !hv || !hc || nv <= 0 || nc <= 0) return I2_E_BADARG;
    mg->prepared = false;
    long long pairs[3] = {0, 0, 0};
    std::vector<long long> pk((size_t)mg->nLocal * 3);
    int rc = for_local(mg, [&](int k) { return host_prepare_mesh(mg->ctx[k], hv, nv, hc, nc, &pk[(size_t)3 * k]); });
    if (rc) return rc;
    for (int c = 0; c < 3; ++c) pairs[c] = pk[c];
    const int W = mg->world;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < W; ++r) {
            mg->lo[c][r] = r == 0 ? 0 : align32(pairs[c] * r / W);
            mg->hi[c][r] = r == W - 1 ? pairs[c] : align32(pairs[c] * (r + 1) / W);
        }
    static const bool equalCuts = [] { const char *e = getenv("I2_MGPU_EQUAL_CUTS"); return e && atoi(e) != 0; }();   // A/B knob
    if (level < 0 && W > 1 && pairs[2] > 0 && !equalCuts) {
        std::vector<double> cost(nc);
        std::vector<unsigned long long> first((size_t)nc + 1);
        rc = i2_host_row_costs(mg->ctx[0], 1, cost.data(), first.data());
        if (rc) return rc;
        double total = 0.0;
        for (int i = 0; i < nc; ++i) total += cost[i];
        std::vector<long long> cut(W + 1, 0);
        cut[W] = pairs[2];
        double run = 0.0;
        int r = 1;
        for (int i = 0; i < nc && r < W; ++i) {
            const double next = run + cost[i];
            while (r < W && next >= total * r / W) {
                const double frac = cost[i] > 0.0 ? (total * r / W - run) / cost[i] : 0.0;   // cut inside the row, linearly
                const long long slot = (long long)first[i] + (long long)(frac * (double)(first[i + 1] - first[i]));
                cut[r] = align32(slot);
                ++r;
            }
            run = next;
        }
        for (; r < W; ++r) cut[r] = align32(pairs[2]);
        for (r = 1; r <= W; ++r) if (cut[r] < cut[r - 1]) cut[r] = cut[r - 1];
        for (r = 0; r < W; ++r) { mg->lo[2][r] = cut[r]; mg->hi[2][r] = cut[r + 1]; }
    }
    rc = for_local(mg, [&](int k) {
        long long lo[3], hi[3];
        for (int c = 0; c < 3; ++c) { lo[c] = mg->lo[c][mg->firstRank + k]; hi[c] = mg->hi[c][mg->firstRank + k]; }
        int e = host_set_forward_ranges(mg->ctx[k], lo, hi);
        if (!e) e = host_prepare_lists(mg->ctx[k]);
        return e;
    });
    if (rc) return rc;
    for (int c = 0; c < 3; ++c) {
        mg->pairs[c] = pairs[c];
        if (taskCounts) taskCounts[c] = 2 * pairs[c];
    }
    mg->nc = nc;
    mg->prepared = true;
    return 0;
}

int i2_mgpu_shard(i2_mgpu *mg, int rank, long long first[3], long long count[3]) {
    if (!mg || !first || !count || rank < 0 || rank >= mg->world) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    for (int c = 0; c < 3; ++c) {
        first[c] = mg->lo[c][rank];
        count[c] = 2 * (mg->hi[c][rank] - mg->lo[c][rank]);
    }
    return 0;
}

int i2_mgpu_set_results_target(i2_mgpu *mg, int local_index, double *const d_results[3]) {
    if (!mg || local_index < 0 || local_index >= mg->nLocal) return I2_E_BADARG;
    for (int c = 0; c < 3; ++c) mg->target[c][local_index] = d_results ? d_results[c] : nullptr;
    return 0;
}

int i2_mgpu_reserve(i2_mgpu *mg, int level, int check) {
    if (!mg) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    return for_local(mg, [&](int k) { return i2_host_reserve(mg->ctx[k], level, check); });
}

// One pass of the hot path over every shard.  Nothing is synchronised unless h_stats is given.
int i2_mgpu_run(i2_mgpu *mg, int level, int check, i2_stats h_stats[3]) {
    if (!mg) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    if (h_stats) std::memset(h_stats, 0, 3 * sizeof(i2_stats));
    mg->lastLevel = level;
    for (int k = 0; k < mg->nLocal; ++k)
        for (int c = 0; c < 3; ++c) mg->ctx[k]->hResultsTarget[c] = mg->target[c][k];
    int rc = for_local(mg, [&](int k) { return host_run_rounds(mg->ctx[k], level); });
    if (rc) return rc;
    NcclApi *api = mg->world > 1 ? nccl_api() : nullptr;
    if (level < 0 && mg->world > 1) {
        if (!api) return I2_E_NCCL;
        I2_NCCL(api->GroupStart());
        for (int k = 0; k < mg->nLocal; ++k) {
            i2_context *c = mg->ctx[k];
            for (int cls = 0; cls < 3; ++cls)
                I2_NCCL(api->AllReduce(&c->scr[cls].qs->lastRound, &c->scr[cls].qs->lastRound, 1, ncclInt32, ncclMax, mg->comm[k], c->stream));
            I2_NCCL(api->AllReduce(c->hRefAll, c->hRefAll, (size_t)3 * c->nc, ncclUint8, ncclMax, mg->comm[k], c->stream));
        }
        I2_NCCL(api->GroupEnd());
    }
    rc = for_local(mg, [&](int k) { return host_run_finalize(mg->ctx[k], level, check != 0); });
    if (rc || !h_stats) return rc;

    // statistics: per-round counts summed over every shard (the lines the reference prints while it iterates)
    long long sums[3][8];   // [class]: unconverged[1..5], orientation warnings, tasks, last round (max)
    std::memset(sums, 0, sizeof(sums));
    for (int k = 0; k < mg->nLocal; ++k) {
        i2_context *c = mg->ctx[k];
        I2_CUDA(cudaSetDevice(c->device));
        QueueState q[3];
        for (int cls = 0; cls < 3; ++cls) {
            std::memset(&q[cls], 0, sizeof(QueueState));
            if (c->hN[cls] > 0) I2_CUDA(cudaMemcpyAsync(&q[cls], c->scr[cls].qs, sizeof(QueueState), cudaMemcpyDeviceToHost, c->stream));
        }
        I2_CUDA(cudaStreamSynchronize(c->stream));
        for (int cls = 0; cls < 3; ++cls) {
            for (int m = 1; m <= MAX_REFINE_LEVEL; ++m) sums[cls][m - 1] += q[cls].count[m];
            sums[cls][5] += q[cls].orientationWarnings;
            sums[cls][6] += c->hN[cls];
            if (q[cls].lastRound > sums[cls][7]) sums[cls][7] = q[cls].lastRound;
        }
    }
    if (mg->world > mg->nLocal) {   // one process per GPU: sum over the processes (24 values through device scratch)
        i2_context *c = mg->ctx[0];
        double v[24];
        for (int cls = 0; cls < 3; ++cls)
            for (int e = 0; e < 8; ++e) v[8 * cls + e] = e == 7 ? 0.0 : (double)sums[cls][e];   // last round: already agreed (adaptive) or 0
        I2_CUDA(cudaMemcpyAsync(mg->scratch[0], v, sizeof(v), cudaMemcpyHostToDevice, c->stream));
        I2_NCCL(api->AllReduce(mg->scratch[0], mg->scratch[0], 24, ncclFloat64, ncclSum, mg->comm[0], c->stream));
        I2_CUDA(cudaMemcpyAsync(v, mg->scratch[0], sizeof(v), cudaMemcpyDeviceToHost, c->stream));
        I2_CUDA(cudaStreamSynchronize(c->stream));
        for (int cls = 0; cls < 3; ++cls)
            for (int e = 0; e < 7; ++e) sums[cls][e] = (long long)(v[8 * cls + e] + 0.5);
    }
    for (int cls = 0; cls < 3; ++cls) {
        i2_stats &st = h_stats[cls];
        const long long n = sums[cls][6];
        st.last_round = (int)sums[cls][7];
        st.orientation_warnings = (int)sums[cls][5];
        st.integrated[0] = n << (level > 0 ? 2 * level : 0);
        long long before = n;
        for (int m = 1; m <= st.last_round && m <= MAX_REFINE_LEVEL; ++m) {
            st.integrated[m] = before << (2 * m);
            st.unconverged[m] = sums[cls][m - 1];
            before = sums[cls][m - 1];
        }
    }
    return 0;
}

// ---- the whole operator (all classes) by row blocks: BASELINE.json configs[3] / [4], meshes beyond the N^2-list limit ----
// Rows are cut by predicted cost under error control (k_row_cost summed per row), equally at a fixed level.  Every GPU
// computes the rows of its block (no cross-GPU sums); the blocks are then combined into the full vector on EVERY GPU by one
// NCCL all-reduce of Point3[nc] (the other blocks are zero on each GPU, so the sum is exact) — the exchange a distributed
// matrix-vector product needs before its next iteration.
int i2_mgpu_apply_prepare(i2_mgpu *mg, const double *hv, int nv, const int *hc, int nc, int level, int *h_row_cuts) {
    if (!mg || !hv || !hc || nv <= 0 || nc <= 0) return I2_E_BADARG;
    mg->applyPrepared = false;
    std::vector<long long> pk((size_t)mg->nLocal * 3);
    int rc = for_local(mg, [&](int k) { return host_prepare_mesh(mg->ctx[k], hv, nv, hc, nc, &pk[(size_t)3 * k]); });
    if (rc) return rc;
    const int W = mg->world;
    mg->rowCut.assign(W + 1, 0);
    mg->rowCut[W] = nc;
    if (level < 0 && W > 1) {
        std::vector<double> cost(nc);
        rc = i2_host_row_costs(mg->ctx[0], 0, cost.data(), nullptr);
        if (rc) return rc;
        double total = 0.0;
        for (int i = 0; i < nc; ++i) total += cost[i];
        double run = 0.0;
        int r = 1;
        for (int i = 0; i < nc && r < W; ++i) {
            run += cost[i];
            while (r < W && run >= total * r / W) mg->rowCut[r++] = i + 1;
        }
        for (; r < W; ++r) mg->rowCut[r] = nc;
    } else {
        for (int r = 1; r < W; ++r) mg->rowCut[r] = (int)((long long)nc * r / W);
    }
    // blocks start at multiples of 32 rows: the warps of a block are warps of the unsharded call (bitwise-equal row sums)
    for (int r = 1; r < W; ++r) {
        mg->rowCut[r] &= ~31;
        if (mg->rowCut[r] < mg->rowCut[r - 1]) mg->rowCut[r] = mg->rowCut[r - 1];
    }
    rc = for_local(mg, [&](int k) {
        i2_context *c = mg->ctx[k];
        const int g = mg->firstRank + k;
        int e = i2_apply_prepare(c, mg->rowCut[g], mg->rowCut[g + 1]);
        if (e) return e;
        if (mg->apFull[k]) cudaFree(mg->apFull[k]);
        if (mg->apW[k]) cudaFree(mg->apW[k]);
        if (mg->apRef[k]) cudaFree(mg->apRef[k]);
        mg->apFull[k] = nullptr; mg->apW[k] = nullptr; mg->apRef[k] = nullptr;
        I2_CUDA(cudaMalloc((void **)&mg->apFull[k], sizeof(double) * 3 * (size_t)nc));
        I2_CUDA(cudaMalloc((void **)&mg->apW[k], sizeof(double) * (size_t)nc));
        I2_CUDA(cudaMalloc((void **)&mg->apRef[k], (size_t)3 * (size_t)(mg->rowCut[g + 1] - mg->rowCut[g]) + 1));
        return 0;
    });
    if (rc) return rc;
    if (h_row_cuts)
        for (int r = 0; r <= W; ++r) h_row_cuts[r] = mg->rowCut[r];
    mg->nc = nc;
    mg->applyPrepared = true;
    return 0;
}

// one application: h_weights double[nc] or NULL; h_out Point3[nc] or NULL (the full vector is then only left on the GPUs, see
// i2_mgpu_apply_result); h_stats[3] or NULL.  Synchronises only when h_out or h_stats is given.
int i2_mgpu_apply(i2_mgpu *mg, int level, const double *h_weights, double *h_out, i2_stats h_stats[3]) {
    if (!mg) return I2_E_BADARG;
    if (!mg->applyPrepared) return I2_E_NOMESH;
    if (level > 0) return I2_E_LEVEL;
    if (h_stats) std::memset(h_stats, 0, 3 * sizeof(i2_stats));
    const int nc = mg->nc;
    NcclApi *api = mg->world > 1 ? nccl_api() : nullptr;
    if (mg->world > 1 && !api) return I2_E_NCCL;
    int rc = for_local(mg, [&](int k) {
        i2_context *c = mg->ctx[k];
        I2_CUDA(cudaSetDevice(c->device));
        if (h_weights) I2_CUDA(cudaMemcpyAsync(mg->apW[k], h_weights, sizeof(double) * nc, cudaMemcpyHostToDevice, c->stream));
        if (mg->world > 1) I2_CUDA(cudaMemsetAsync(mg->apFull[k], 0, sizeof(double) * 3 * (size_t)nc, c->stream));
        return i2_apply_rounds(c, level, h_weights ? mg->apW[k] : nullptr);
    });
    if (rc) return rc;
    if (level < 0 && mg->world > 1) {
        I2_NCCL(api->GroupStart());
        for (int k = 0; k < mg->nLocal; ++k) {
            i2_context *c = mg->ctx[k];
            int *L[3] = {&c->scr[0].qs->lastRound, &c->scr[1].qs->lastRound, reinterpret_cast<int *>(c->partial + 6)};
            for (int cls = 0; cls < 3; ++cls) I2_NCCL(api->AllReduce(L[cls], L[cls], 1, ncclInt32, ncclMax, mg->comm[k], c->stream));
        }
        I2_NCCL(api->GroupEnd());
    }
    std::vector<i2_stats> part((size_t)3 * mg->nLocal);
    rc = for_local(mg, [&](int k) {
        i2_context *c = mg->ctx[k];
        const int g = mg->firstRank + k;
        return i2_apply_finish(c, level, h_weights ? mg->apW[k] : nullptr, mg->apFull[k] + 3 * (size_t)mg->rowCut[g], mg->apRef[k],
                               h_stats ? &part[(size_t)3 * k] : nullptr);
    });
    if (rc) return rc;
    if (mg->world > 1) {
        I2_NCCL(api->GroupStart());
        for (int k = 0; k < mg->nLocal; ++k)
            I2_NCCL(api->AllReduce(mg->apFull[k], mg->apFull[k], (size_t)3 * nc, ncclFloat64, ncclSum, mg->comm[k], mg->ctx[k]->stream));
        I2_NCCL(api->GroupEnd());
    }
    if (h_out) {
        i2_context *c = mg->ctx[0];
        I2_CUDA(cudaSetDevice(c->device));
        I2_CUDA(cudaMemcpyAsync(h_out, mg->apFull[0], sizeof(double) * 3 * (size_t)nc, cudaMemcpyDeviceToHost, c->stream));
        I2_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (h_stats) {
        double v[3][14];   // per class: integrated[0], unconverged[1..5], warnings, last round (max: already agreed)
        std::memset(v, 0, sizeof(v));
        for (int k = 0; k < mg->nLocal; ++k)
            for (int cls = 0; cls < 3; ++cls) {
                const i2_stats &p = part[(size_t)3 * k + cls];
                v[cls][0] += (double)p.integrated[0];
                for (int m = 1; m <= MAX_REFINE_LEVEL; ++m) v[cls][m] += (double)p.unconverged[m];
                v[cls][6] += p.orientation_warnings;
                if (p.last_round > v[cls][7]) v[cls][7] = p.last_round;
            }
        if (mg->world > mg->nLocal) {
            i2_context *c = mg->ctx[0];
            double flat[24];
            for (int cls = 0; cls < 3; ++cls)
                for (int e = 0; e < 8; ++e) flat[8 * cls + e] = e == 7 ? 0.0 : v[cls][e];
            I2_CUDA(cudaSetDevice(c->device));
            I2_CUDA(cudaMemcpyAsync(mg->scratch[0], flat, sizeof(flat), cudaMemcpyHostToDevice, c->stream));
            I2_NCCL(api->AllReduce(mg->scratch[0], mg->scratch[0], 24, ncclFloat64, ncclSum, mg->comm[0], c->stream));
            I2_CUDA(cudaMemcpyAsync(flat, mg->scratch[0], sizeof(flat), cudaMemcpyDeviceToHost, c->stream));
            I2_CUDA(cudaStreamSynchronize(c->stream));
            for (int cls = 0; cls < 3; ++cls)
                for (int e = 0; e < 7; ++e) v[cls][e] = flat[8 * cls + e];
        }
        for (int cls = 0; cls < 3; ++cls) {
            i2_stats &st = h_stats[cls];
            st.last_round = (int)v[cls][7];
            st.orientation_warnings = (int)(v[cls][6] + 0.5);
            st.integrated[0] = (long long)(v[cls][0] + 0.5);
            long long before = st.integrated[0];
            for (int m = 1; m <= st.last_round && m <= MAX_REFINE_LEVEL; ++m) {
                st.integrated[m] = before << (2 * m);
                st.unconverged[m] = (long long)(v[cls][m] + 0.5);
                before = st.unconverged[m];
            }
        }
    }
    return 0;
}

// device pointer of the full result vector Point3[nc] on local GPU k (valid after i2_mgpu_apply + a stream synchronisation) and
// of the block's per-class refinement counters unsigned char[3][rows of the block]
int i2_mgpu_apply_result(i2_mgpu *mg, int local_index, double **d_full, unsigned char **d_refinements) {
    if (!mg || local_index < 0 || local_index >= mg->nLocal) return I2_E_BADARG;
    if (!mg->applyPrepared) return I2_E_NOMESH;
    if (d_full) *d_full = mg->apFull[local_index];
    if (d_refinements) *d_refinements = mg->apRef[local_index];
    return 0;
}

int i2_mgpu_checksums(i2_mgpu *mg, double sums[12]) {
    if (!mg || !sums) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    for (int e = 0; e < 12; ++e) sums[e] = 0.0;
    if (mg->world > mg->nLocal) {
        // one process per GPU: checksum kernels, all-reduce on the device, ONE copy back and one synchronisation
        NcclApi *api = nccl_api();
        if (!api) return I2_E_NCCL;
        i2_context *c = mg->ctx[0];
        I2_CUDA(cudaSetDevice(c->device));
        double *d = mg->scratch[0];
        I2_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 12, c->stream));
        for (int k = 0; k < 3; ++k)
            launch_checksum(c->hResultsTarget[k] ? c->hResultsTarget[k] : c->hResults[k], c->hN[k], d + 4 * k, c->numSMs, c->stream);
        I2_CUDA(cudaGetLastError());
        I2_NCCL(api->AllReduce(d, d, 12, ncclFloat64, ncclSum, mg->comm[0], c->stream));
        I2_CUDA(cudaMemcpyAsync(sums, d, sizeof(double) * 12, cudaMemcpyDeviceToHost, c->stream));
        I2_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    }
    for (int k = 0; k < mg->nLocal; ++k) {
        double part[12];
        const int rc = i2_host_checksums(mg->ctx[k], part);
        if (rc) return rc;
        for (int e = 0; e < 12; ++e) sums[e] += part[e];
    }
    return 0;
}

// Export gather: the result shards of one class to GPU `root` (global rank), concatenated in rank order — the row of
// rank r starts at task 2 * lo_r, the order of the shard-local lists.  what = 0: results (Point3), 1: tasks (int3), 2: defects.
// d_dst is read on the process that owns `root` only.  Enqueued on the contexts' streams.
int i2_mgpu_gather(i2_mgpu *mg, int cls, int what, int root, void *d_dst) {
    if (!mg || cls < 0 || cls > 2 || what < 0 || what > 2 || root < 0 || root >= mg->world) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    const bool rootLocal = root >= mg->firstRank && root < mg->firstRank + mg->nLocal;
    if (rootLocal && !d_dst) return I2_E_BADARG;
    const size_t elem = what == 0 ? sizeof(double) * 3 : (what == 1 ? sizeof(int) * 3 : sizeof(double));
    for (int k = 0; k < mg->nLocal && what == 2; ++k)
        if (mg->ctx[k]->hN[cls] > 0 && !mg->ctx[k]->hErrors[cls]) return I2_E_BADARG;   // the last run did not compute the defects
    auto src = [&](int k) -> const char * {
        i2_context *c = mg->ctx[k];
        if (what == 2) return (const char *)c->hErrors[cls];
        return what == 0 ? (const char *)(c->hResultsTarget[cls] ? c->hResultsTarget[cls] : c->hResults[cls]) : (const char *)c->hTasks[cls];
    };
    NcclApi *api = mg->world > 1 ? nccl_api() : nullptr;
    if (mg->world > 1 && !api) return I2_E_NCCL;
    if (api) I2_NCCL(api->GroupStart());
    for (int k = 0; k < mg->nLocal; ++k) {
        i2_context *c = mg->ctx[k];
        const int g = mg->firstRank + k;
        I2_CUDA(cudaSetDevice(c->device));
        if (g == root) {
            for (int r = 0; r < mg->world; ++r) {
                const long long n = 2 * (mg->hi[cls][r] - mg->lo[cls][r]);
                char *at = (char *)d_dst + elem * (size_t)(2 * mg->lo[cls][r]);
                if (n == 0) continue;
                if (r == root) I2_CUDA(cudaMemcpyAsync(at, src(k), elem * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
                else I2_NCCL(api->Recv(at, elem * (size_t)n, ncclUint8, r, mg->comm[k], c->stream));
            }
        } else {
            const long long n = c->hN[cls];
            if (n > 0) I2_NCCL(api->Send(src(k), elem * (size_t)n, ncclUint8, root, mg->comm[k], c->stream));
        }
    }
    if (api) I2_NCCL(api->GroupEnd());
    return 0;
}

// Row-striped export: the shard of local GPU `local_index` to host arrays (shard-sized, the shard's own order; any may be NULL)
int i2_mgpu_fetch(i2_mgpu *mg, int local_index, int cls, int *h_tasks, double *h_results, double *h_errors) {
    if (!mg || local_index < 0 || local_index >= mg->nLocal) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    return i2_host_fetch(mg->ctx[local_index], cls, h_tasks, h_results, h_errors);
}

// (max, mean) of the (i,j)/(j,i) defects of a class over all shards (the last run must have been called with check != 0)
int i2_mgpu_error_summary(i2_mgpu *mg, int cls, double out[2]) {
    if (!mg || cls < 0 || cls > 2 || !out) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    double mx = 0.0, sum = 0.0, cnt = 0.0;
    for (int k = 0; k < mg->nLocal; ++k) {
        i2_context *c = mg->ctx[k];
        if (c->hN[cls] == 0) continue;
        if (!c->hErrors[cls]) return I2_E_BADARG;
        double part[2];
        const int rc = i2_error_summary(c, c->hErrors[cls], c->hN[cls], part);
        if (rc) return rc;
        if (part[0] > mx) mx = part[0];
        sum += part[1] * (double)c->hN[cls];
        cnt += (double)c->hN[cls];
    }
    if (mg->world > mg->nLocal) {
        NcclApi *api = nccl_api();
        if (!api) return I2_E_NCCL;
        i2_context *c = mg->ctx[0];
        I2_CUDA(cudaSetDevice(c->device));
        double v[3] = {sum, cnt, mx};
        I2_CUDA(cudaMemcpyAsync(mg->scratch[0], v, sizeof(v), cudaMemcpyHostToDevice, c->stream));
        I2_NCCL(api->AllReduce(mg->scratch[0], mg->scratch[0], 2, ncclFloat64, ncclSum, mg->comm[0], c->stream));
        I2_NCCL(api->AllReduce(mg->scratch[0] + 2, mg->scratch[0] + 2, 1, ncclFloat64, ncclMax, mg->comm[0], c->stream));
        I2_CUDA(cudaMemcpyAsync(v, mg->scratch[0], sizeof(v), cudaMemcpyDeviceToHost, c->stream));
        I2_CUDA(cudaStreamSynchronize(c->stream));
        sum = v[0]; cnt = v[1]; mx = v[2];
    }
    out[0] = mx;
    out[1] = cnt > 0.0 ? sum / cnt : 0.0;
    return 0;
}

int i2_mgpu_refinements(i2_mgpu *mg, int cls, unsigned char *h_refinements) {
    if (!mg || cls < 0 || cls > 2 || !h_refinements) return I2_E_BADARG;
    if (!mg->prepared) return I2_E_NOMESH;
    i2_context *c = mg->ctx[0];
    I2_CUDA(cudaSetDevice(c->device));
    I2_CUDA(cudaMemcpyAsync(h_refinements, c->hRefinements[cls], c->nc, cudaMemcpyDeviceToHost, c->stream));
    I2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
