// Kernel-side declarations shared by i2_kernels.cu (device code) and i2_abi.cu (C-ABI entry points).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include "i2_vec.cuh"

namespace i2 {

// Packed triangle data in HBM: structure of arrays, component-major with a padded stride so that
// consecutive triangles are consecutive in memory (coalesced when consecutive lanes hold consecutive j).
// See DESIGN.md "Data layout in HBM".
enum PackComp {
    PK_A = 0,    // 0..2   vertex A
    PK_B = 3,    // 3..5   vertex B
    PK_C = 6,    // 6..8   vertex C
    PK_TA = 9,   // 9..11  unit tangent (C-B)^
    PK_TB = 12,  // 12..14 unit tangent (A-C)^
    PK_TC = 15,  // 15..17 unit tangent (B-A)^
    PK_NU = 18,  // 18..20 (B-A)x(C-A), not normalised
    PK_N = 21,   // 21..23 unit normal n_j (as Mesh3D computes it)
    PK_S = 24,   // 24     area S
    PK_L = 25,   // 25..27 edge lengths |C-B|, |A-C|, |B-A|
    PK_COUNT = 28
};

struct PackedMesh {
    const double *tri;   // [PK_COUNT][stride]
    const int *cells;    // int3[nc] (vertex ids, needed to locate shared vertices/edges)
    int nc;
    int stride;
};

// device-resident bookkeeping of the adaptive work queue (one per context)
struct QueueState {
    int count[MAX_REFINE_LEVEL + 2];   // count[m] = tasks still unconverged after round m (count[0] = all)
    int lastRound;                     // L = last refinement round that was executed
    int orientationWarnings;
};

extern std::atomic<long long> g_launchCount;       // kernels launched by the launch_* wrappers (bench.py's gpu_launches)
constexpr int kThreads = 128;          // CTA size of the integrate kernels

// math mode of the regular-pair point function
enum MathMode { MATH_STRICT = 0, MATH_FAST = 1, MATH_FAST_LIBDEVICE = 2, MATH_FAST_POINTWISE = 3 };

void launch_pack(const double *verts, const int *cells, const double *normals, const double *measures, int nc, int stride,
                 double *tri, cudaStream_t s);
void launch_geometry(const double *verts, const int *cells, int nc, double *normals, double *centers, double *measures,
                     cudaStream_t s);
// regular part of `count` tasks at uniform refinement `level`; list==nullptr -> task slots 0..count-1,
// otherwise slots list[0..*countDev-1] (device-side count, persistent grid)
// half > 0 (only with list == nullptr): the tasks are two segments [0, half) and [half, count) whose warp groups are formed
// independently (pairs / reversed pairs of runAllPairs), see k_regular_grouped
void launch_integrate(int cls, int mathMode, const PackedMesh &pm, const int *tasks, const int *list, const int *countDev,
                      long long countHost, long long half, int level, double *out4, double *fusedResults3, int numSMs, cudaStream_t s, int flags = 0);
void launch_apply_regular(const PackedMesh &pm, int rowLo, int rowHi, int colLo, int colHi, int chunks, const double *weights,
                          double *partial, double *out3, cudaStream_t s);
// list-free regular class with the Runge loop per pair (see k_apply_regular_adaptive)
void launch_apply_regular_adaptive(const PackedMesh &pm, int rowLo, int rowHi, int colLo, int colHi, int chunks, const double *weights,
                                   double *partial6, unsigned char *depth, int *lastRound, unsigned long long *counts6, double *out3,
                                   double *other3, unsigned char *refinements, cudaStream_t s);
void launch_reduce_partials_adaptive(const double *partial6, const unsigned char *depth, int rows, int chunks, const int *lastRound, double *out3,
                                     double *other3, unsigned char *refinements, cudaStream_t s);
void launch_checksum(const double *results3, long long n, double *sums4, int numSMs, cudaStream_t s);
// Runge comparison of round `round` + deterministic compaction of the unconverged slots: staging = int[capacity of the list]
// (per-CTA segments), blockCnt = int[kCompareMaxBlocks], listOut = dense list in input order, *countOut = its length
constexpr int kCompareMaxBlocks = 148 * 32;
void launch_compare(const double *cur4, const double *prev4, const int *tasks, const int *listIn, const int *countIn,
                    long long countHost, int *staging, int *blockCnt, int *listOut, int *countOut, unsigned char *cellFlag,
                    unsigned char *converged, QueueState *qs, int round, int numSMs, cudaStream_t s);
void launch_flag_cells(const int *tasks, long long n, unsigned char *cellFlag, cudaStream_t s);
void launch_bump(unsigned char *cellFlag, unsigned char *refinements, int nc, cudaStream_t s);
// adds the closed-form singular integral (adjacent classes), assembles J; bufA/bufB selected by QueueState::lastRound
void launch_finalize(int cls, const PackedMesh &pm, const double *verts, const int *tasks, long long n, double *bufA4,
                     const double *bufB4, const QueueState *qs, double *results3, QueueState *qsMut, cudaStream_t s);
void launch_symmetry_error(const double *results3, long long nHalf, double *errors, cudaStream_t s);
void launch_add_reversed(int *tasks3, long long n, cudaStream_t s);
void launch_max_vertex_id(const int *cells, int nc, int *maxId, cudaStream_t s);
// out2[0] = max, out2[1] = sum of n non-negative doubles (the (i,j)/(j,i) defects): the --checkresults summary
void launch_error_summary(const double *errors, long long n, double *out2, int numSMs, cudaStream_t s);
void launch_split_uniform(const double *vin, int nvIn, const int *cin, int ncIn, const double *min, double *vout, int *cout,
                          double *mout, cudaStream_t s);
// classification by vertex incidence + sharded task lists (i2_prepare.cu)
size_t incidence_scratch_ints(int nv, int nc);
// incidence, per-row partner counts, rowOff[3][nc + 1] (first slot of every row per class, last entry = total), totals[3]
void launch_incidence(const int *cells, int nv, int nc, int *scratch, unsigned long long *rowOff, unsigned long long *totals, cudaStream_t s);
void launch_partners_fill(const int *cells, int nv, int nc, const int *scratch, const unsigned long long *rowOff, int *simple, int *attached,
                          bool reversed, cudaStream_t s);
void launch_regular_fill(const int *cells, int nc, const unsigned long long *rowOff, unsigned long long fLo, unsigned long long fHi, int *out,
                         int *outRev, int numSMs, cudaStream_t s);
void launch_row_cost(const PackedMesh &pm, bool upperOnly, double *cost, cudaStream_t s);
// row-major adjacent lists for the operator apply (all partners j != i of the rows of a block, sorted by (i, j))
size_t rows_scratch_ints(int nc);
void launch_partners_both(const int *cells, int nv, int nc, const int *scratch, int *scratch2, cudaStream_t s);
void launch_partners_fill_rows(const int *cells, int nv, int nc, const int *scratch, const int *scratch2, int rowLo, int rowHi, int *simple,
                               int *attached, cudaStream_t s);
void launch_row_scatter(const int *tasks, const double *results, const int *off, int rowLo, int rows, const double *weights, double *out,
                        cudaStream_t s);
void launch_take_rows(const unsigned char *perCell, int rowLo, int rows, unsigned char *out, cudaStream_t s);
void launch_selftest_math(int op, const double *a, const double *b, long long n, double *out, cudaStream_t s);
cudaError_t upload_math_tables(cudaStream_t s);
cudaError_t preload_kernels();
cudaError_t upload_quadrature(const double *Lxyzw, int n, double pow2p, cudaStream_t s, bool *shape13 = nullptr);
// FP64-pipe / MUFU peak micro-benchmarks: returns elapsed ms for `iters` dependent-chain iterations
void launch_peak_dfma(double *sink, int iters, int blocks, cudaStream_t s);
void launch_peak_dfma3(double *sink, const double *seed, int iters, int blocks, cudaStream_t s);
void launch_peak_mufu(double *sink, int iters, int blocks, cudaStream_t s);
void launch_peak_mix(int ni, double *sink, int iters, int blocks, cudaStream_t s);

}  // namespace i2
