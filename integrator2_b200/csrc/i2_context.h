// Internal: the context behind the opaque i2_context of include/i2_abi.h, shared by i2_abi.cu and i2_mgpu.cu.
#pragma once
#include "../../include/i2_abi.h"
#include "i2_kernels.cuh"

#include <cstddef>

struct i2_context {
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;
    bool ownStream = false;
    int mathMode = I2_MATH_FAST;
    bool haveQuad = false;
    bool ruleShape13 = false;   // the rule has the 1 + 3 + 3 + 6 equal-weight structure of Cowper's 13-point rule (straight-line kernel variant)

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
        i2::QueueState *qs = nullptr;
    } scr[3];
    // i2_integrate_all / i2_host_run: the two adjacent classes run on side streams next to the regular class
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t forkEv = nullptr, sideDone[2] = {nullptr, nullptr};

    // matrix-free path scratch (per-chunk partial row sums)
    double *partial = nullptr;
    size_t partialCap = 0;
    unsigned char *depthBuf = nullptr;   // list-free adaptive path: per (column chunk, row) refinement depth
    size_t depthCap = 0;

    // classification scratch of i2_classify_count / _fill (Mesh3D's lists): first slot of every row per class [3][nc + 1], totals ...
    unsigned long long *rowCounts = nullptr;
    size_t rowCap = 0;
    int rowCountsNc = -1, rowCountsNv = 0; // mesh of the last i2_classify_count (i2_classify_fill must see the same one)
    int *clsScratch = nullptr;            // its vertex incidence
    size_t clsScratchCap = 0;
    // ... and of the vertex-incidence path (i2_host_prepare): CSR scratch, first slot of every row per class, totals
    int *incScratch = nullptr;
    size_t incCap = 0;
    unsigned long long *rowOff = nullptr;   // [3][nc + 1] + totals[3]
    size_t rowOffCap = 0;
    double *rowCost = nullptr;              // predicted adaptive cost per row (i2_host_row_costs)
    size_t rowCostCap = 0;

    // host-entry state (i2_host_prepare / i2_host_run)
    double *hVerts = nullptr, *hNormals = nullptr, *hMeasures = nullptr;
    int *hCells = nullptr;
    int *hTasks[3] = {nullptr, nullptr, nullptr};
    double *hIntegrals[3] = {nullptr, nullptr, nullptr};
    double *hResults[3] = {nullptr, nullptr, nullptr};
    double *hResultsTarget[3] = {nullptr, nullptr, nullptr};   // when set: where the class's results go instead (e.g. peer-mapped export array)
    double *hErrors[3] = {nullptr, nullptr, nullptr};
    unsigned char *hRefinements[3] = {nullptr, nullptr, nullptr};
    long long hCount[3] = {0, 0, 0};       // ordered tasks of the WHOLE mesh per class (2 x pairs)
    int hNv = 0;
    bool hPrepared = false;                 // i2_host_prepare completed for the mesh in hVerts/hCells
    // multi-GPU use of the host-entry path: this context owns, per class, the pairs with forward slots [hLo, hLo + hHalf) in
    // BOTH orders; its task list is [those pairs ; their reversed pairs] (hN = 2 hHalf tasks), i.e. a small runAllPairs list
    int shardRank = 0, shardWorld = 1;
    bool explicitRanges = false;
    long long rangeLo[3] = {0, 0, 0}, rangeHi[3] = {0, 0, 0};
    long long hLo[3] = {0, 0, 0}, hHalf[3] = {0, 0, 0}, hN[3] = {0, 0, 0};
    unsigned char *hRefAll = nullptr;       // hRefinements[k] = hRefAll + k * nc (one allocation: one all-reduce in multi-GPU runs)
    size_t capRefAll = 0;
    int *adjFull[2] = {nullptr, nullptr};   // whole ordered lists of the two adjacent classes (shards are copied out of them)
    size_t capAdjFull[2] = {0, 0};
    bool runFused = false;                  // last host_run_rounds fused the regular class's assembly
    // the ~70 dependent launches of the three classes' error-control chains, captured once per prepared shard and replayed
    // (CUDA graph): the counts live on the device and the sequence is fixed, only the launch latency is at stake
    cudaGraphExec_t roundsGraph = nullptr;
    unsigned long long roundsGraphKey = 0;
    long long roundsGraphLaunches = 0;      // kernels per replay (for i2_launch_count)
    size_t capVerts = 0, capCells = 0, capNormals = 0, capMeasures = 0, capTasks[3] = {0, 0, 0}, capIntegrals[3] = {0, 0, 0},
           capResults[3] = {0, 0, 0}, capErrors[3] = {0, 0, 0};
    double *sumScratch = nullptr;           // 16 doubles: checksums / summaries without a cudaMalloc per call
    bool incidenceValid = false;            // incScratch / rowOff describe the mesh currently set
    // operator apply (i2_apply_*): row block, row-major adjacent lists of the block, their buffers
    struct ApplyState {
        bool prepared = false;
        int rowLo = 0, rowHi = 0, chunks = 0;
        int *scratch2 = nullptr;            // both-sided partner counts and offsets (rows_scratch_ints)
        size_t scratch2Cap = 0;
        int *tasks[2] = {nullptr, nullptr};
        double *integrals[2] = {nullptr, nullptr}, *results[2] = {nullptr, nullptr};
        size_t capTasks[2] = {0, 0}, capIntegrals[2] = {0, 0}, capResults[2] = {0, 0};
        long long n[2] = {0, 0};
        unsigned char *refCells = nullptr;  // [2][nc] per-cell refinement counters of the adjacent classes
        size_t refCellsCap = 0;
        double *regular = nullptr;          // fixed level: row sums of the regular class before the adjacent ones are added
        size_t regularCap = 0;
    } ap;
    cudaEvent_t chunkDone[2] = {nullptr, nullptr};
    bool profiling = false;
    cudaEvent_t prof[3] = {nullptr, nullptr, nullptr};
};


namespace i2 {
// Internal (non-ABI) pieces of i2_host_run that the multi-GPU layer interleaves with its collectives (i2_mgpu.cu):
//   host_run_rounds   : integration rounds of the three classes (fork/join of the side streams included); at a fixed level the
//                       regular class is complete afterwards (fused assembly)
//   host_run_finalize : closed-form singular parts + final assembly (reads each class's QueueState::lastRound, which a multi-GPU
//                       caller has replaced by the maximum over all ranks), optional (i,j)/(j,i) defect
int host_prepare_mesh(i2_context *c, const double *hv, int nv, const int *hc, int nc, long long pairs[3]);
int host_prepare_lists(i2_context *c);
int host_run_rounds(i2_context *c, int level);
int ensure_incidence(i2_context *c);   // vertex incidence + per-row offsets for the mesh currently set (no-op when valid)
int host_run_finalize(i2_context *c, int level, bool wantErrors);
// explicit forward-slot ranges of this context's shard (cost-balanced multi-GPU runs); lo/hi are rounded down to multiples of 32
int host_set_forward_ranges(i2_context *c, const long long lo[3], const long long hi[3]);
}  // namespace i2
