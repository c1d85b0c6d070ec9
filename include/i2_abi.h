/*
 * i2_abi.h — C ABI of the B200-native integrator2 hot path (libintegrator2_b200.so).
 *
 * The reference (andreyypopov/integrator2) has no C ABI: its hot path sits behind the C++ abstract class
 * Evaluator3D (src/evaluators/evaluator3d.cuh:51-53, pure virtuals integrateOverSimpleNeighbors /
 * integrateOverAttachedNeighbors / integrateOverNotNeighbors) implemented by EvaluatorJ3DK
 * (src/evaluators/evaluatorJ3DK.cu:849-1012).  The functions below are what a C/FFI binding of that path binds;
 * the drop-in C++ classes in include/integrator2/ are thin callers of exactly these entry points
 * (see INTEGRATION.md for the mapping and for a ctypes / cgo style stub).
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer on the context's device, `h_` = host pointer;
 *   - Point3 arrays are double[n][3], double4 arrays are double[n][4] (32-byte aligned), int3 arrays int[n][3]:
 *     the same memory layouts as the reference's Point3/double4/int3 device vectors;
 *   - neighbour classes use the reference's enum values (src/Mesh3d.cuh:18-23):
 *     0 = simple (vertex-adjacent), 1 = attached (edge-adjacent), 2 = not neighbours (regular);
 *   - every function returns 0 on success, a positive cudaError_t, or a negative I2_E_* code;
 *   - all work of a context is enqueued on one CUDA stream; functions taking h_ outputs synchronise it;
 *   - there is no CPU fallback: without a CUDA device every call fails with a CUDA error code.
 */
#ifndef I2_ABI_H
#define I2_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct i2_context i2_context;

#define I2_E_BADARG (-1)     /* null pointer / negative size / class out of range                      */
#define I2_E_NOMESH (-2)     /* i2_set_mesh was not called                                            */
#define I2_E_NOQUAD (-3)     /* i2_set_quadrature was not called                                      */
#define I2_E_LEVEL (-4)      /* fixed refinement level outside 0..12                                  */
#define I2_E_TOOBIG (-5)     /* a count does not fit the 32-bit slots of the reference's int3 tasks   */
#define I2_E_NCCL (-6)       /* multi-GPU layer: libnccl.so.2 could not be loaded or an NCCL call failed */

#define I2_CLASS_SIMPLE 0
#define I2_CLASS_ATTACHED 1
#define I2_CLASS_NOT 2

#define I2_LEVEL_ADAPTIVE (-1) /* automatic error control (Runge rule), NumericalIntegrator3D default     */

#define I2_MATH_STRICT 0 /* regular pairs evaluated in the reference's operation order                    */
#define I2_MATH_FAST 1   /* hoisted / un-normalised formulation, grouped logs/atan2, branch-free primitives (default) */
#define I2_MATH_FAST_LIBDEVICE 2 /* hoisted algebra, point by point, libdevice sqrt/log/atan2 (diagnostic)  */
#define I2_MATH_FAST_POINTWISE 3 /* hoisted algebra, point by point, branch-free primitives (diagnostic)    */

/* per-class statistics of one i2_integrate_class call (printed by the drop-in classes exactly like the
 * reference prints them: src/evaluators/evaluatorJ3DK.cu:956,980 and src/evaluators/evaluator3d.cu:338) */
typedef struct i2_stats {
    int last_round;              /* L: number of refinement rounds executed (0 in fixed mode)                */
    long long integrated[6];     /* refined tasks integrated in round r = unconverged(r-1) * 4^r             */
    long long unconverged[6];    /* original tasks still unconverged after round r (r >= 1)                  */
    int orientation_warnings;    /* coplanar vertex-adjacent pairs with opposite normals                     */
} i2_stats;

/* ---- context ---------------------------------------------------------------------------------------- */
int i2_create(i2_context **ctx, int device);
int i2_destroy(i2_context *ctx);
int i2_set_stream(i2_context *ctx, void *cuda_stream);   /* optional: run on the caller's stream            */
int i2_synchronize(i2_context *ctx);
int i2_set_math_mode(i2_context *ctx, int mode);
const char *i2_error_string(int code);

/* ---- quadrature: replaces NumericalIntegrator3D's constant-memory upload
 *      (src/NumericalIntegrator3d.cu:197-211).  xy = n pairs (L_x, L_y), L_z = 1 - L_x - L_y;
 *      order = p of the Runge rule (2^p).  Process-global constant memory, like the reference.            */
int i2_set_quadrature(i2_context *ctx, const double *h_xy, const double *h_w, int n, int order);

/* ---- mesh: replaces kCalculateCellNormal/Center/Measure (src/Mesh3d.cu:23-71); any output may be NULL   */
int i2_mesh_geometry(i2_context *ctx, const double *d_vertices, int nv, const int *d_cells, int nc,
                     double *d_normals, double *d_centers, double *d_measures);
/* packs vertices/tangents/normals/areas of all triangles into the SoA the kernels read (context-owned);
 * the four input arrays are borrowed and must stay alive while the context uses the mesh                  */
int i2_set_mesh(i2_context *ctx, const double *d_vertices, int nv, const int *d_cells, int nc,
                const double *d_normals, const double *d_measures);

/* one uniform midpoint refinement of a whole mesh, materialised (only the refined-mesh EXPORT needs it; the
 * integrate kernels rebuild children on the fly).  Replaces kSplitCell for the fixed-level case
 * (src/NumericalIntegrator3d.cu:37-87) with deterministic slots: the 3 new vertices of cell c go to
 * nv_in + 3c.., its 4 children to 4c..4c+3.  Outputs: vertices[nv_in + 3 nc_in], cells[4 nc_in], measures[4 nc_in] */
int i2_refine_mesh_once(i2_context *ctx, const double *d_vertices_in, int nv_in, const int *d_cells_in, int nc_in,
                        const double *d_measures_in, double *d_vertices_out, int *d_cells_out, double *d_measures_out);

/* ---- neighbour classification: replaces kDetermineNeighborType (src/Mesh3d.cu:93-142, 243-262).
 *      By vertex incidence: the partners of a triangle are the triangles listed under its three vertices (once = one shared
 *      vertex, twice = an edge), O(N valence^2); the regular class is what remains of each row and is only enumerated when its
 *      list is asked for.  Count, then fill, so that the caller sizes its lists exactly; the lists come out deterministic
 *      (lexicographic in (i,j), i<j, k = slot) where the reference's order is atomicAdd order.                     */
int i2_classify_count(i2_context *ctx, const int *d_cells, int nc, long long h_counts[3]);
int i2_classify_fill(i2_context *ctx, const int *d_cells, int nc, int *d_simple, int *d_attached, int *d_not);
/* any of the three lists may be NULL and is then skipped: a mesh whose regular list would not fit (N^2/2 x 12 B) takes
 * only the two adjacent lists from here and integrates the regular class list-free (i2_apply_regular[_adaptive])       */
/* tasks[n..2n) = (j, i, n+idx): replaces kAddReversedPairs (src/evaluators/evaluator3d.cu:22-31)            */
int i2_add_reversed_pairs(i2_context *ctx, int *d_tasks, long long n);

/* ---- THE HOT PATH: replaces EvaluatorJ3DK::integrateOver{Simple,Attached,Not}Neighbors
 *      (src/evaluators/evaluatorJ3DK.cu:849-1012) including numericalIntegration, the refinement loop of
 *      NumericalIntegrator3D (src/NumericalIntegrator3d.cu:369-558) and compareIntegrationResults
 *      (src/evaluators/evaluator3d.cu:297-341).
 *   d_tasks      int3[n]  (i, j, k); the result of task t is stored at slot t
 *   level        >= 0 fixed uniform refinement of the control panel i, I2_LEVEL_ADAPTIVE for error control
 *   d_integrals  double4[n] out: (Psi, Theta) including the closed-form singular part
 *   d_results    Point3[n]  out: J(K_i, K_j)
 *   d_refinements unsigned char[nc] in/out, adaptive only, may be NULL: += 1 per round in which cell i still
 *                has an unconverged task (NumericalIntegrator3D::getRefinementsRequired)
 *   d_converged  unsigned char[n] out, adaptive only, may be NULL (getIntegralsConverged)
 *   h_stats      may be NULL; when given the call synchronises the stream before returning                 */
int i2_integrate_class(i2_context *ctx, int cls, const int *d_tasks, long long n, int level,
                       double *d_integrals, double *d_results, unsigned char *d_refinements,
                       unsigned char *d_converged, i2_stats *h_stats);

/* The same for the ordered list Evaluator3D::runAllPairs builds (src/evaluators/evaluator3d.cu:156-169): d_tasks = n_half pairs
 * (i, j) followed by their n_half reversed pairs (j, i), 2 n_half tasks in all.  The regular-pair kernel takes a few far-field
 * decisions per warp of 32 consecutive tasks; here the warps of the reversed half start at its first task, so a result depends
 * only on the task's position inside its half: a GPU that integrates the pairs [lo, hi) and their reversed pairs, lo and hi
 * multiples of 32 (i2_host_set_shard / i2_mgpu_*), produces the very bits of the unsharded call.                              */
int i2_integrate_pairs(i2_context *ctx, int cls, const int *d_tasks, long long n_half, int level,
                       double *d_integrals, double *d_results, unsigned char *d_refinements,
                       unsigned char *d_converged, i2_stats *h_stats);

/* The three classes of Evaluator3D::runAllPairs (src/evaluators/evaluator3d.cu:120-204 calls the three virtuals one
 * after the other) in ONE call: arrays indexed by I2_CLASS_*; same arguments and results as three i2_integrate_class
 * calls, but the two adjacent classes (short, latency-bound chains of kernels) are enqueued on internal side streams and
 * overlap the regular class; the context's stream waits for them before the call returns (fork/join with events, no
 * host synchronisation unless h_stats is given).  n[k] = 0 skips class k; d_refinements / d_converged / h_stats may be
 * NULL, and so may their elements.                                                                               */
int i2_integrate_all(i2_context *ctx, const int *const d_tasks[3], const long long n[3], int level,
                     double *const d_integrals[3], double *const d_results[3], unsigned char *const d_refinements[3],
                     unsigned char *const d_converged[3], i2_stats h_stats[3]);

/* ---- list-free regular class ("next" row f1 of SURVEY.md §8: implicit tiled enumeration, no N^2 list) ---------------
 * d_out[i - row_lo] = sum over all triangles j that share no vertex with i (j != i) of w_j * J(K_i, K_j), level 0,
 * for rows row_lo <= i < row_hi; d_weights = double[nc] or NULL (all ones); d_out = Point3[row_hi - row_lo].
 * The pairs are enumerated on the fly (the vertex-id comparison of src/Mesh3d.cu:115-122 is the classification), no
 * task list and no per-pair result is stored: this is the entry point for meshes beyond the reference's int32 / N^2
 * limits (N > 46 340 triangles) and the unit that shards by rows across GPUs.                                    */
int i2_apply_regular(i2_context *ctx, int row_lo, int row_hi, const double *d_weights, double *d_out);

/* The same row sums under AUTOMATIC ERROR CONTROL: every pair (i, j) runs the Runge loop of
 * EvaluatorJ3DK::numericalIntegration (src/evaluators/evaluatorJ3DK.cu:895-1012) — round 0 on the control panel, round 1 on
 * its 4 children, criterion of src/evaluators/evaluator3d.cu:76-99, further rounds (16 ... 1024 children, at most 5) only
 * for the pairs that fail it — inside one kernel, without task lists, refined meshes or a work queue in memory.  The value
 * of a pair follows the reference's ping-pong buffers: the newest value of the buffer selected by the parity
 * of the class's last round L.  d_out = sums for this call's own L (h_stats->last_round); d_out_other (may be NULL) = sums
 * for the other parity, which a multi-GPU caller takes when the maximum of L over all ranks has the other parity.
 * d_refinements (may be NULL) = unsigned char[row_hi - row_lo], the per-cell value NumericalIntegrator3D::
 * getRefinementsRequired(not_neighbors) would hold: 1 + the number of compare rounds some pair of the row failed.
 * h_stats (may be NULL; synchronises): integrated[0] = pairs examined, unconverged[m] = pairs failing round m.        */
int i2_apply_regular_adaptive(i2_context *ctx, int row_lo, int row_hi, const double *d_weights, double *d_out,
                              double *d_out_other, unsigned char *d_refinements, i2_stats *h_stats);

/* ---- the whole operator for a block of rows, all neighbour classes: d_out[i - row_lo] = sum over j != i of w_j J(K_i, K_j) ----
 * The regular class runs list-free (i2_apply_regular at level 0, i2_apply_regular_adaptive under error control); the vertex- and
 * edge-adjacent partners of the rows come from the vertex incidence as row-major task lists (all partners j of row i, sorted by
 * (i, j)), are integrated by the same kernels as i2_integrate_class — regular part by quadrature, closed-form singular part
 * (src/evaluators/evaluatorJ3DK.cu:849-881) — and added to the row sums in list order (deterministic).  This is the entry point
 * for meshes beyond the reference's N^2-list limit (BASELINE.json configs[3], [4]) and shards by rows: no cross-GPU sums.
 *   i2_apply_prepare   rows [row_lo, row_hi) of the mesh given to i2_set_mesh (or uploaded by i2_host_prepare / i2_mgpu_*)
 *   i2_apply           level = 0 or I2_LEVEL_ADAPTIVE; d_weights double[nc] or NULL; d_out Point3[rows];
 *                      d_refinements unsigned char[3][rows] or NULL (error control: the per-cell counters of the three classes
 *                      for the block's rows); h_stats[3] or NULL (synchronises)
 *   i2_apply_rounds / i2_apply_last_rounds / i2_apply_finish   the same in two halves for multi-GPU callers: under error control
 *                      the GPUs agree on each class's last round (maximum; h_last = {simple, attached, regular}) in between */
int i2_apply_prepare(i2_context *ctx, int row_lo, int row_hi);
int i2_apply(i2_context *ctx, int level, const double *d_weights, double *d_out, unsigned char *d_refinements, i2_stats h_stats[3]);
int i2_apply_rounds(i2_context *ctx, int level, const double *d_weights);
int i2_apply_last_rounds(i2_context *ctx, int h_last[3], int set);
int i2_apply_finish(i2_context *ctx, int level, const double *d_weights, double *d_out, unsigned char *d_refinements, i2_stats h_stats[3]);

/* delta = |J_ij + J_ji|_1 / max(|J_ij|_1, |J_ji|_1) for slots t and n_half+t: replaces
 * kCalculateIntegrationError (src/evaluators/evaluator3d.cu:45-57)                                         */
int i2_symmetry_error(i2_context *ctx, const double *d_results, long long n_half, double *d_errors);

/* h_out = (max, mean) of n non-negative device doubles: the summary of the (i,j)/(j,i) defects that --checkresults computes
 * but the reference never prints (src/evaluators/evaluator3d.cu:188-203)                                       */
int i2_error_summary(i2_context *ctx, const double *d_errors, long long n, double h_out[2]);

/* ---- host-buffer entry points (the end-to-end path: host mesh in, host results out) --------------------
 * i2_host_prepare uploads the mesh (already scaled), computes geometry, classifies and builds the three
 * ordered task lists exactly like Evaluator3D::runAllPairs (src/evaluators/evaluator3d.cu:120-169);
 * h_task_counts[c] = 2 * pairs of class c.  i2_host_run integrates all three classes and copies tasks,
 * results (and deltas when h_errors[c] != NULL) back; copies of finished chunks overlap the computation
 * of the next ones when the host buffers are pinned.  Any h_tasks[c]/h_results[c] may be NULL (skipped).
 * The classification is by vertex incidence (partners of a triangle = the triangles listed under its three vertices), the
 * regular list is implicit in per-row prefix sums and is written together with its reversed pairs by one kernel.         */
int i2_host_prepare(i2_context *ctx, const double *h_vertices, int nv, const int *h_cells, int nc,
                    long long h_task_counts[3]);
int i2_host_run(i2_context *ctx, int level, int *const h_tasks[3], double *const h_results[3],
                double *const h_errors[3], unsigned char *const h_refinements[3], i2_stats h_stats[3]);
/* per class (sum J_x, sum J_y, sum J_z, sum |J|_1) of the results left in device memory by i2_host_run: the small
 * 'metric' a caller reads back when the per-pair results stay resident (what Evaluator3D::runAllPairs leaves behind) */
int i2_host_checksums(i2_context *ctx, double h_sums[12]);
/* multi-GPU use of the host-buffer path (one context per GPU; i2_mgpu_* below drives it): rank r of w owns, per class, the
 * pairs whose forward slots lie in [lo_r, hi_r) — equal counts, lo and hi multiples of 32 — in BOTH orders, and its task list is
 * [those pairs ; their reversed pairs]: the shape of a small runAllPairs list, so the (i,j)/(j,i) defect is local and, because
 * warp groups are formed per half from multiples of 32, every result has the same bits as in the unsharded run.  Call
 * i2_host_set_shard before i2_host_prepare (default: 0 of 1 = everything); i2_host_prepare still returns the FULL counts but
 * materialises only the shard (classification by vertex incidence: no O(N^2) pass, the regular list is filled for the shard's
 * slot range alone); i2_host_shard returns lo_r and the shard's task count 2 (hi_r - lo_r) per class; the host buffers given to
 * i2_host_run are shard-sized and i2_host_checksums covers the shard.                                                    */
int i2_host_set_shard(i2_context *ctx, int rank, int world);
int i2_host_shard(i2_context *ctx, long long h_first[3], long long h_count[3]);
/* predicted cost of every row's regular pairs under error control (level-0 pair integrations: 5 + 66.7 P(rho), rho = centroid
 * distance / sqrt(larger area)), upper_only: pairs j > i counted twice; and the first regular forward slot of every row
 * (nc + 1 values).  Either output may be NULL.  Used to place cost-balanced shard cuts (i2_mgpu_prepare, level < 0).      */
int i2_host_row_costs(i2_context *ctx, int upper_only, double *h_cost, unsigned long long *h_row_first_regular);
/* The two halves of i2_host_run for callers that shard over several contexts with their OWN communicator (MPI, gloo, ...):
 * i2_host_run_rounds enqueues the integration rounds of the shard; under error control the shards must then agree on each class's
 * last round (maximum; i2_host_last_rounds gets / sets the three device-side values) and on the per-cell refinement counters
 * (element-wise maximum; i2_host_refinements gets / sets unsigned char[3][nc]); i2_host_run_finalize adds the closed-form
 * singular parts, assembles J and, with check != 0, computes the (i,j)/(j,i) defects; i2_host_fetch copies a class of the shard
 * (tasks, results, defects; any may be NULL) to the host.  i2_mgpu_run is exactly this sequence with NCCL in the middle.     */
/* allocates the scratch the next run of this level needs (work queue, second result buffer, defect arrays): keeps cudaMalloc out
 * of a caller's timed region (the reference allocates its buffers before its timers start, src/evaluators/evaluator3d.cu:122-154) */
int i2_host_reserve(i2_context *ctx, int level, int check);
int i2_host_run_rounds(i2_context *ctx, int level);
int i2_host_last_rounds(i2_context *ctx, int h_last[3], int set);
int i2_host_refinements(i2_context *ctx, unsigned char *h_refinements, int set);
int i2_host_run_finalize(i2_context *ctx, int level, int check);
int i2_host_fetch(i2_context *ctx, int cls, int *h_tasks, double *h_results, double *h_errors);
/* device views of what i2_host_prepare built (valid until the next prepare/destroy)                        */
int i2_host_device_views(i2_context *ctx, const int *d_tasks[3], const double *d_results[3]);

/* ---- multi-GPU: the pair lists of Evaluator3D::runAllPairs (src/evaluators/evaluator3d.cu:120-204) sharded over the GPUs of
 * one box (no reference counterpart: /root/reference is single-GPU; SURVEY.md §8(b),(e)).  One handle drives the GPUs of this
 * process: all of them from one process (i2_mgpu_create_local: CLI / host classes, env I2_GPUS) or one per process
 * (i2_mgpu_create_rank; rank 0 calls i2_mgpu_unique_id and ships the 128 bytes to the others by any means).  NCCL is loaded
 * with dlopen("libnccl.so.2") on first use; world = 1 needs no NCCL.
 *   i2_mgpu_prepare   host mesh in; every GPU uploads it, classifies by vertex incidence and builds ITS shard of the three
 *                     ordered lists (see i2_host_set_shard); level < 0 places the cuts of the regular class by predicted
 *                     adaptive cost.  h_task_counts = ordered tasks of the whole mesh.
 *   i2_mgpu_run       one pass of the hot path over every shard (level >= 0 or I2_LEVEL_ADAPTIVE; check != 0 also computes the
 *                     (i,j)/(j,i) defects).  Error control: NCCL all-reduce(max) of each class's last round before the final
 *                     assembly (the value of a converged pair depends on the parity of the GLOBAL last round) and
 *                     of the per-cell refinement counters.  Asynchronous unless h_stats (per-round counts summed over all
 *                     shards) is given.
 *                     With one process per GPU every call that returns job-wide numbers is COLLECTIVE — i2_mgpu_run with h_stats,
 *                     i2_mgpu_checksums, i2_mgpu_error_summary, i2_mgpu_apply with h_stats, and i2_mgpu_run / i2_mgpu_apply under
 *                     error control in any case: all ranks make the same calls in the same order.
 *   i2_mgpu_checksums per class (sum J_x, J_y, J_z, sum |J|_1) over ALL shards (all-reduce), the step's small metric
 *   i2_mgpu_shard     forward-slot range of any rank: first slot and task count 2 (hi - lo) per class
 *   i2_mgpu_fetch     row-striped export: one local GPU's shard (tasks, results, defects; shard order) to host arrays
 *   i2_mgpu_gather    export to one GPU: result (what = 0), task (1) or defect (2) shards of a class concatenated in rank order into
 *                     d_dst on GPU `root` (ncclSend / ncclRecv); enqueued on the contexts' streams
 *   i2_mgpu_set_results_target  make local GPU k write class c's results to d_results[c] instead of its own buffer, e.g. into a
 *                     peer-mapped slice of the exporting GPU's array (i2_peer_*): compute and gather in one kernel           */
typedef struct i2_mgpu i2_mgpu;
#define I2_MGPU_ID_BYTES 128
int i2_mgpu_unique_id(unsigned char h_id[I2_MGPU_ID_BYTES]);
int i2_mgpu_create_rank(i2_mgpu **mg, int device, int rank, int world, const unsigned char h_id[I2_MGPU_ID_BYTES]);
int i2_mgpu_create_local(i2_mgpu **mg, int ngpus, const int *devices /* NULL: 0 .. ngpus-1 */);
int i2_mgpu_destroy(i2_mgpu *mg);
int i2_mgpu_info(i2_mgpu *mg, int *world, int *n_local, int *first_rank);
i2_context *i2_mgpu_context(i2_mgpu *mg, int local_index);
int i2_mgpu_set_quadrature(i2_mgpu *mg, const double *h_xy, const double *h_w, int n, int order);
int i2_mgpu_set_math_mode(i2_mgpu *mg, int mode);
int i2_mgpu_synchronize(i2_mgpu *mg);
int i2_mgpu_prepare(i2_mgpu *mg, const double *h_vertices, int nv, const int *h_cells, int nc, int level, long long h_task_counts[3]);
int i2_mgpu_shard(i2_mgpu *mg, int rank, long long h_first[3], long long h_count[3]);
int i2_mgpu_set_results_target(i2_mgpu *mg, int local_index, double *const d_results[3]);
int i2_mgpu_reserve(i2_mgpu *mg, int level, int check);   /* i2_host_reserve on every local GPU */
int i2_mgpu_run(i2_mgpu *mg, int level, int check, i2_stats h_stats[3]);
int i2_mgpu_checksums(i2_mgpu *mg, double h_sums[12]);
int i2_mgpu_gather(i2_mgpu *mg, int cls, int what, int root, void *d_dst);
int i2_mgpu_fetch(i2_mgpu *mg, int local_index, int cls, int *h_tasks, double *h_results, double *h_errors);
int i2_mgpu_refinements(i2_mgpu *mg, int cls, unsigned char *h_refinements);
int i2_mgpu_error_summary(i2_mgpu *mg, int cls, double h_out[2]);   /* (max, mean) defect of a class over all shards */
/* the whole operator (i2_apply_*) by row blocks over the GPUs: BASELINE.json configs[3] / [4], meshes beyond the N^2-list limit.
 * i2_mgpu_apply_prepare uploads the mesh and cuts the rows — by predicted cost under error control (level < 0), equally
 * otherwise; h_row_cuts (int[world + 1]) may be NULL.  i2_mgpu_apply: every GPU computes its rows (regular class list-free,
 * adjacent classes from row-major lists; under error control one all-reduce(max) of the three last rounds in between), then ONE
 * NCCL all-reduce of Point3[nc] leaves the full vector on every GPU.  h_weights double[nc] or NULL, h_out Point3[nc] or NULL,
 * h_stats[3] or NULL (counts summed over the GPUs); asynchronous when both h_out and h_stats are NULL.                      */
int i2_mgpu_apply_prepare(i2_mgpu *mg, const double *h_vertices, int nv, const int *h_cells, int nc, int level, int *h_row_cuts);
int i2_mgpu_apply(i2_mgpu *mg, int level, const double *h_weights, double *h_out, i2_stats h_stats[3]);
int i2_mgpu_apply_result(i2_mgpu *mg, int local_index, double **d_full, unsigned char **d_refinements);

/* ---- multi-GPU export by peer stores: per-pair results written straight into the exporting GPU's memory over NVLink --
 * One process per GPU.  The exporting rank
 * allocates the full result array with i2_peer_alloc (cudaMalloc + CUDA IPC handle, 64 opaque bytes the host side ships
 * to the other processes by any means, e.g. torch.distributed.broadcast_object_list); every other rank maps it with
 * i2_peer_open and passes `mapped + 24 * first_task_of_its_shard` as d_results of i2_integrate_class / i2_integrate_all or to
 * i2_mgpu_set_results_target:
 * the final-assembly stores of the kernels then ARE the gather (16-byte coalesced warp stores through NVSwitch, no
 * staging copy, no second pass), overlapped with the arithmetic of the other warps.  The data are visible to the owner
 * once the writer's stream has been synchronised (i2_synchronize) and the processes have met at a host barrier.
 * i2_peer_close unmaps (writer side), i2_peer_free releases (owner side).                                           */
#define I2_PEER_HANDLE_BYTES 64
int i2_peer_alloc(i2_context *ctx, unsigned long long bytes, void **d_ptr, unsigned char h_handle[I2_PEER_HANDLE_BYTES]);
int i2_peer_open(i2_context *ctx, const unsigned char h_handle[I2_PEER_HANDLE_BYTES], void **d_ptr);
int i2_peer_close(i2_context *ctx, void *d_ptr);
int i2_peer_free(i2_context *ctx, void *d_ptr);

/* ---- instrumentation for bench.py -------------------------------------------------------------------------
 * launch counter: kernels launched by this library since process start (all contexts).
 * profiling: when enabled, i2_integrate_class brackets its integrate kernel(s) and its finalize kernel with CUDA
 * events on the context's stream; i2_profile_last returns the device times of the last fixed-level call.      */
int i2_launch_count(long long *h_count);
int i2_set_profiling(i2_context *ctx, int enabled);
int i2_profile_last(i2_context *ctx, float *ms_integrate, float *ms_finalize);

/* self-test hook: evaluates one of the device math primitives of the regular-pair kernel element-wise on device arrays
 * (op 0: fast_sqrt(a), 1: fast_rcp(a), 2: log_ratio(a, b), 3: atan2_fast(a, b)); used by tests/test_math_primitives.py */
int i2_selftest_math(i2_context *ctx, int op, const double *d_a, const double *d_b, long long n, double *d_out);

/* ---- measured roofline denominators: FP64-pipe DFMA rate and XU-pipe MUFU rate of this device ---------- */
int i2_peak_rates(i2_context *ctx, double *dfma_tflops, double *mufu_gops);
/* DFMA rate when every instruction reads three distinct 64-bit register operands (register-file bandwidth included) */
int i2_peak_dfma_three_operand(i2_context *ctx, double *tflops);
/* DFMA rate with int_per_dfma (0..3) independent integer instructions interleaved per DFMA: tells whether the dispatch port is
 * free during the second cycle of a warp-wide FP64 instruction (rate unchanged) or not (rate drops)                          */
int i2_peak_dfma_with_integer(i2_context *ctx, int int_per_dfma, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* I2_ABI_H */
