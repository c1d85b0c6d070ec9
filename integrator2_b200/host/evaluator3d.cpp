// Evaluator3D host class: task lists, result buffers, the per-class driver and the exports.
#include "../../include/integrator2/evaluators/evaluator3d.cuh"
#include "host_context.h"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

Evaluator3D::Evaluator3D(const Mesh3D &mesh_, NumericalIntegrator3D &numIntegrator_) : mesh(mesh_), numIntegrator(numIntegrator_) {}

namespace {
struct ClassBuffers {
    deviceVector<int3> *tasks;
    deviceVector<double4> *integrals;
    deviceVector<Point3> *results;
    deviceVector<double> *errors;
};
}  // namespace

#define I2_CLASS_BUFFERS(self)                                                                                             \
    ClassBuffers cb[3] = {{&(self)->simpleNeighborsTasks, &(self)->d_simpleNeighborsIntegrals, &(self)->d_simpleNeighborsResults, &(self)->simpleNeighborsErrors},     \
                          {&(self)->attachedNeighborsTasks, &(self)->d_attachedNeighborsIntegrals, &(self)->d_attachedNeighborsResults, &(self)->attachedNeighborsErrors}, \
                          {&(self)->notNeighborsTasks, &(self)->d_notNeighborsIntegrals, &(self)->d_notNeighborsResults, &(self)->notNeighborsErrors}}

const deviceVector<Point3> *Evaluator3D::getResultsVector(neighbour_type_enum t) const {
    auto *self = const_cast<Evaluator3D *>(this);
    I2_CLASS_BUFFERS(self);
    return ((int)t >= 0 && (int)t < 3) ? cb[(int)t].results : nullptr;
}
const deviceVector<double4> *Evaluator3D::getIntegralsVector(neighbour_type_enum t) const {
    auto *self = const_cast<Evaluator3D *>(this);
    I2_CLASS_BUFFERS(self);
    return ((int)t >= 0 && (int)t < 3) ? cb[(int)t].integrals : nullptr;
}
const deviceVector<double> *Evaluator3D::getErrorsVector(neighbour_type_enum t) const {
    auto *self = const_cast<Evaluator3D *>(this);
    I2_CLASS_BUFFERS(self);
    return ((int)t >= 0 && (int)t < 3) ? cb[(int)t].errors : nullptr;
}

void Evaluator3D::allocateClass(int cls, int taskCount) {
    I2_CLASS_BUFFERS(this);
    cb[cls].tasks->allocate(taskCount);
    cb[cls].results->allocate(taskCount);
    cb[cls].integrals->allocate(taskCount);
}

void Evaluator3D::reportDefects(int cls, double maxDelta, double meanDelta, long long n) {
    // the reference computes the (i,j)/(j,i) defects under --checkresults but prints nothing (they only appear as the Error
    // column of the exports); one summary line per class makes the check usable without an export
    static const char *names[3] = {"simple neighbors", "attached neighbors", "not neighbors"};
    summary[cls].deltaMax = maxDelta;
    summary[cls].deltaMean = meanDelta;
    printf("Symmetry check (i,j)/(j,i) for %s: max delta = %g, mean delta = %g (%lld ordered pairs)\n", names[cls], maxDelta, meanDelta, n);
}

// Same sequence as the reference's runAllPairs (src/evaluators/evaluator3d.cu:120-204): ordered tasks = pairs followed by
// the reversed pairs, preparation, the three timed per-class integrations, optional (i,j)/(j,i) defect.
void Evaluator3D::runAllPairs(bool checkCorrectness) {
    for (auto &cs : summary) cs = ClassSummary();
    lastGpus = i2host::gpus();
    const bool adaptive = numIntegrator.getErrorControlType() == error_control_type_enum::automatic_error_control;
    lastLevel = adaptive ? I2_LEVEL_ADAPTIVE : numIntegrator.getFixedRefinementLevel();
    if (lastGpus > 1) {
        runAllPairsMultiGpu(checkCorrectness);
        return;
    }
    distributed = false;
    i2_context *ctx = i2host::context();
    I2_CLASS_BUFFERS(this);
    const deviceVector<int3> *pairs[3] = {&mesh.getSimpleNeighbors(), &mesh.getAttachedNeighbors(), &mesh.getNotNeighbors()};
    for (int k = 0; k < 3; ++k) {
        allocateClass(k, 2 * pairs[k]->size);
        copy_d2d(pairs[k]->data, cb[k].tasks->data, pairs[k]->size);
        checkI2Errors(i2_add_reversed_pairs(ctx, (int *)cb[k].tasks->data, pairs[k]->size));
        summary[k].tasks = 2LL * pairs[k]->size;
    }
    numIntegrator.prepareTasksAndMesh(simpleNeighborsTasks, attachedNeighborsTasks, notNeighborsTasks);
    tasksArePairs = true;

    nvtxRangePushA("Simple neighbors integration");
    timer.start();
    integrateOverSimpleNeighbors();
    summary[0].ms = timer.stop("Simple neighbors integration");
    nvtxRangePop();
    requestFreeDeviceMemoryAmount();

    nvtxRangePushA("Attached neighbors integration");
    timer.start();
    integrateOverAttachedNeighbors();
    summary[1].ms = timer.stop("Attached neighbors integration");
    nvtxRangePop();
    requestFreeDeviceMemoryAmount();

    nvtxRangePushA("Non-neighbors integration");
    timer.start();
    integrateOverNotNeighbors();
    summary[2].ms = timer.stop("Non-neighbors integration");
    nvtxRangePop();
    requestFreeDeviceMemoryAmount();
    tasksArePairs = false;

    if (checkCorrectness) {
        nvtxRangePushA("(i,j)/(j,i) defects");
        for (int k = 0; k < 3; ++k) {
            cb[k].errors->allocate(2 * pairs[k]->size);
            checkI2Errors(i2_symmetry_error(ctx, (const double *)cb[k].results->data, pairs[k]->size, cb[k].errors->data));
        }
        checkCudaErrors(cudaDeviceSynchronize());
        for (int k = 0; k < 3; ++k) {
            double mm[2] = {0.0, 0.0};
            checkI2Errors(i2_error_summary(ctx, cb[k].errors->data, 2LL * pairs[k]->size, mm));
            reportDefects(k, mm[0], mm[1], 2LL * pairs[k]->size);
        }
        nvtxRangePop();
    }
}

// The same run sharded over the GPUs of the box (env I2_GPUS > 1) through the multi-GPU layer of the C ABI: every GPU builds
// and integrates its own shard of the three ordered lists; under error control the GPUs agree on the last rounds and the
// per-cell refinement counters with NCCL inside i2_mgpu_run.  The per-pair results stay where they were computed (row-striped);
// outputResultsToFile collects them from all GPUs and writes the file the single-GPU run writes; I2_GATHER=1 additionally
// collects tasks, results and defects in this object's device vectors on device 0 (shards concatenated in rank order).
void Evaluator3D::runAllPairsMultiGpu(bool checkCorrectness) {
    i2_mgpu *mg = i2host::mgpu();
    I2_CLASS_BUFFERS(this);
    const auto &hv = mesh.getHostVertices();
    const auto &hc = mesh.getHostCells();
    long long counts[3] = {0, 0, 0};
    nvtxRangePushA("sharded prepare");
    checkI2Errors(i2_mgpu_prepare(mg, (const double *)hv.data(), (int)hv.size(), (const int *)hc.data(), (int)hc.size(), lastLevel, counts));
    nvtxRangePop();
    // mode bookkeeping of the integrator (refinement counters, refined mesh for the export) with lists that live elsewhere
    deviceVector<int3> sized[3];
    for (int k = 0; k < 3; ++k) {
        sized[k].size = (int)counts[k];
        distributedCount[k] = counts[k];
        summary[k].tasks = counts[k];
    }
    numIntegrator.prepareTasksAndMesh(sized[0], sized[1], sized[2]);
    for (int k = 0; k < 3; ++k) sized[k].size = 0;
    printf("\nIntegrating over simple neighbors (%lld pairs), attached neighbors (%lld pairs) and not neighbors (%lld pairs) on %d GPUs...\n",
           counts[0], counts[1], counts[2], lastGpus);

    checkI2Errors(i2_mgpu_reserve(mg, lastLevel, checkCorrectness ? 1 : 0));   // buffers before the timer, like the reference
    checkI2Errors(i2_mgpu_synchronize(mg));
    i2_stats st[3];
    nvtxRangePushA("all classes integration (multi-GPU)");
    const auto t0 = std::chrono::steady_clock::now();
    checkI2Errors(i2_mgpu_run(mg, lastLevel, checkCorrectness ? 1 : 0, st));
    checkI2Errors(i2_mgpu_synchronize(mg));
    allClassesMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    nvtxRangePop();
    static const char *names[3] = {"simple neighbors", "attached neighbors", "not neighbors"};
    for (int k = 0; k < 3; ++k) {
        summary[k].lastRound = st[k].last_round;
        for (int m = 0; m < 6; ++m) summary[k].unconverged[m] = st[k].unconverged[m];
        if (lastLevel < 0) {
            printf("%s:\n", names[k]);
            printf("Iteration 0, integrating %lld tasks\n", counts[k]);
            long long checked = counts[k];
            for (int m = 1; m <= st[k].last_round; ++m) {
                printf("Iteration %d, integrating %lld tasks\n", m, st[k].integrated[m]);
                printf("Out of %lld tasks: %lld converged, %lld did not converge\n", checked, checked - st[k].unconverged[m], st[k].unconverged[m]);
                checked = st[k].unconverged[m];
            }
        }
    }
    printf("Time for Simple neighbors + Attached neighbors + Non-neighbors integration on %d GPUs: %6.3f ms\n", lastGpus, allClassesMs);
    distributed = true;
    distributedChecked = checkCorrectness;
    if (lastLevel < 0) {
        std::vector<unsigned char> ref(hc.size());
        for (int k = 0; k < 3; ++k) {
            checkI2Errors(i2_mgpu_refinements(mg, k, ref.data()));
            auto *dst = numIntegrator.getRefinementsRequired(neighbour_type_enum(k));
            if (dst && dst->data) copy_h2d(ref.data(), dst->data, (int)ref.size());
        }
        checkCudaErrors(cudaDeviceSynchronize());
    }
    if (checkCorrectness)
        for (int k = 0; k < 3; ++k) {
            double mm[2] = {0.0, 0.0};
            checkI2Errors(i2_mgpu_error_summary(mg, k, mm));
            reportDefects(k, mm[0], mm[1], counts[k]);
        }
    if (i2host::gatherToDevice0()) {
        for (int k = 0; k < 3; ++k) {
            if (!counts[k]) continue;
            cb[k].tasks->allocate((int)counts[k]);
            cb[k].results->allocate((int)counts[k]);
            checkI2Errors(i2_mgpu_gather(mg, k, 1, 0, cb[k].tasks->data));
            checkI2Errors(i2_mgpu_gather(mg, k, 0, 0, cb[k].results->data));
            if (checkCorrectness) {
                cb[k].errors->allocate((int)counts[k]);
                checkI2Errors(i2_mgpu_gather(mg, k, 2, 0, cb[k].errors->data));
            }
        }
        checkI2Errors(i2_mgpu_synchronize(mg));
    } else {
        fprintf(stderr, "integrator2 (B200 build): I2_GPUS=%d — per-pair results stay row-striped over the GPUs (outputResultsToFile collects them); "
                        "set I2_GATHER=1 to also gather them into the evaluator's device vectors on device 0\n", lastGpus);
    }
}

// User-supplied task lists (src/evaluators/evaluator3d.cu:213-288): empty classes are skipped, no reversed pairs, no defect.
void Evaluator3D::runPairs(const std::vector<int3> &userSimple, const std::vector<int3> &userAttached, const std::vector<int3> &userNot) {
    if (userSimple.empty() && userAttached.empty() && userNot.empty()) return;
    I2_CLASS_BUFFERS(this);
    const std::vector<int3> *user[3] = {&userSimple, &userAttached, &userNot};
    for (int k = 0; k < 3; ++k) {
        if (user[k]->empty()) continue;
        allocateClass(k, (int)user[k]->size());
        copy_h2d(user[k]->data(), cb[k].tasks->data, user[k]->size());
    }
    checkCudaErrors(cudaDeviceSynchronize());
    numIntegrator.prepareTasksAndMesh(simpleNeighborsTasks, attachedNeighborsTasks, notNeighborsTasks);
    tasksArePairs = false;
    distributed = false;

    if (!userSimple.empty()) {
        timer.start();
        integrateOverSimpleNeighbors();
        timer.stop("Simple neighbors integration");
        requestFreeDeviceMemoryAmount();
    }
    if (!userAttached.empty()) {
        timer.start();
        integrateOverAttachedNeighbors();
        timer.stop("Attached neighbors integration");
        requestFreeDeviceMemoryAmount();
    }
    if (!userNot.empty()) {
        timer.start();
        integrateOverNotNeighbors();
        timer.stop("Non-neighbors integration");
        requestFreeDeviceMemoryAmount();
    }
}

int Evaluator3D::compareIntegrationResults(neighbour_type_enum, bool) {
    static bool said = false;
    if (!said) {
        said = true;
        fprintf(stderr, "integrator2 (B200 build): Evaluator3D::compareIntegrationResults does nothing here — the Runge comparison and the compaction of "
                        "unconverged tasks run on the device inside i2_integrate_class (include/i2_abi.h)\n");
    }
    return 0;
}

// Export format of the reference (src/evaluators/evaluator3d.cu:354-457): default ostream precision,
// csv rows "i;j;x;y;z[;err]" with a quoted header, plain text rows "(i, j): [x, y, z][, error = e]".
// output_format_enum::binary (not in the reference, whose exports keep 6 significant digits): "<Name>.bin" =
//   char[4] "I2RB", int32 version = 1, int64 n, int32 hasErrors, int32 reserved, then int32 tasks[n][3] (i, j, k),
//   float64 results[n][3], float64 errors[n] when hasErrors — full precision, keyed (i, j), little endian.
bool Evaluator3D::outputResultsToFile(neighbour_type_enum neighborType, output_format_enum outputFormat) const {
    const int k = (int)neighborType;
    if (k < 0 || k > 2) return false;
    auto *self = const_cast<Evaluator3D *>(this);
    I2_CLASS_BUFFERS(self);
    const int n = distributed ? (int)distributedCount[k] : cb[k].tasks->size;
    const char *ext = outputFormat == output_format_enum::csv ? ".csv" : (outputFormat == output_format_enum::binary ? ".bin" : ".dat");
    std::string filename = neighborTypeString(neighborType) + ext;
    if (!n) return false;

    std::vector<Point3> hostResults(n);
    std::vector<int3> hostTasks(n);
    std::vector<double> hostErrors;
    bool withErrors;
    nvtxRangePushA("export: device to host");
    if (distributed) {
        // row-striped: every GPU delivers the rows it computed, [pairs lo..hi ; their reversed pairs]; they are placed where the
        // single-GPU list has them (all pairs first, then all reversed pairs), so the file does not depend on the number of GPUs
        i2_mgpu *mg = i2host::mgpu();
        withErrors = distributedChecked;
        if (withErrors) hostErrors.resize(n);
        int world = 1;
        checkI2Errors(i2_mgpu_info(mg, &world, nullptr, nullptr));
        const long long P = n / 2;
        std::vector<std::thread> th;
        std::vector<int> rcs(world, 0);
        for (int r = 0; r < world; ++r)
            th.emplace_back([&, r] {
                long long first[3], count[3];
                if ((rcs[r] = i2_mgpu_shard(mg, r, first, count))) return;
                const long long m = count[k], h = m / 2, lo = first[k];
                if (!m) return;
                std::vector<int3> t(m);
                std::vector<Point3> res(m);
                std::vector<double> err(withErrors ? m : 0);
                if ((rcs[r] = i2_mgpu_fetch(mg, r, k, (int *)t.data(), (double *)res.data(), withErrors ? err.data() : nullptr))) return;
                std::memcpy(&hostTasks[lo], t.data(), sizeof(int3) * h);
                std::memcpy(&hostTasks[P + lo], t.data() + h, sizeof(int3) * h);
                std::memcpy(&hostResults[lo], res.data(), sizeof(Point3) * h);
                std::memcpy(&hostResults[P + lo], res.data() + h, sizeof(Point3) * h);
                if (withErrors) {
                    std::memcpy(&hostErrors[lo], err.data(), sizeof(double) * h);
                    std::memcpy(&hostErrors[P + lo], err.data() + h, sizeof(double) * h);
                }
            });
        for (auto &t : th) t.join();
        for (int rc : rcs) checkI2Errors(rc);
    } else {
        copy_d2h(cb[k].results->data, hostResults.data(), n);
        copy_d2h(cb[k].tasks->data, hostTasks.data(), n);
        withErrors = cb[k].errors->data != nullptr;
        if (withErrors) {
            hostErrors.resize(n);
            copy_d2h(cb[k].errors->data, hostErrors.data(), n);
        }
        checkCudaErrors(cudaDeviceSynchronize());
    }
    nvtxRangePop();

    FILE *out = fopen(filename.c_str(), outputFormat == output_format_enum::binary ? "wb" : "w");
    if (!out) {
        printf("Error while opening the file\n");
        return false;
    }
    if (outputFormat == output_format_enum::binary) {
        const int version = 1, hasErrors = withErrors ? 1 : 0, reserved = 0;
        const long long n64 = n;
        fwrite("I2RB", 1, 4, out);
        fwrite(&version, sizeof(int), 1, out);
        fwrite(&n64, sizeof(long long), 1, out);
        fwrite(&hasErrors, sizeof(int), 1, out);
        fwrite(&reserved, sizeof(int), 1, out);
        fwrite(hostTasks.data(), sizeof(int3), n, out);
        fwrite(hostResults.data(), sizeof(Point3), n, out);
        if (withErrors) fwrite(hostErrors.data(), sizeof(double), n, out);
        fclose(out);
        printf("%d results saved to file %s\n", n, filename.c_str());
        return true;
    }
    nvtxRangePushA("export: format and write");
    if (outputFormat == output_format_enum::csv) {
        fputs("\"TaskI\";\"TaskJ\";\"IntegralX\";\"IntegralY\";\"IntegralZ\"", out);
        if (withErrors) fputs(";\"Error\"", out);
        fputc('\n', out);
    }
    // The reference streams every row through `ofstream << double << std::endl` (default precision: "%g", 6 digits) and
    // flushes per row; the same bytes are produced here by formatting blocks of rows in parallel into strings that are
    // written in order (2.9e8 rows on Vint16k: the formatting, not the GPU, is the export's cost).
    const int block = 1 << 16;
    const int nBlocks = (n + block - 1) / block;
    const int wave = 64;   // blocks formatted concurrently, then written in order
    std::vector<std::string> text(wave);
    for (int b0 = 0; b0 < nBlocks; b0 += wave) {
        const int nb = std::min(wave, nBlocks - b0);
#pragma omp parallel for schedule(dynamic, 1)
        for (int b = 0; b < nb; ++b) {
            std::string &dst = text[b];
            dst.clear();
            char line[256];
            const int lo = (b0 + b) * block, hi = std::min(n, lo + block);
            for (int t = lo; t < hi; ++t) {
                const int3 task = hostTasks[t];
                const Point3 J = hostResults[t];
                int len;
                if (outputFormat == output_format_enum::csv) {
                    len = snprintf(line, sizeof(line), "%d;%d;%g;%g;%g", task.x, task.y, J.x, J.y, J.z);
                    if (withErrors) len += snprintf(line + len, sizeof(line) - len, ";%g", hostErrors[t]);
                } else {
                    len = snprintf(line, sizeof(line), "(%d, %d): [%g, %g, %g]", task.x, task.y, J.x, J.y, J.z);
                    if (withErrors) len += snprintf(line + len, sizeof(line) - len, ", error = %g", hostErrors[t]);
                }
                line[len++] = '\n';
                dst.append(line, len);
            }
        }
        for (int b = 0; b < nb; ++b) fwrite(text[b].data(), 1, text[b].size(), out);
    }
    fclose(out);
    nvtxRangePop();
    printf("%d results saved to file %s\n", n, filename.c_str());
    return true;
}

std::string Evaluator3D::getRunSummaryJson() const {
    static const char *names[3] = {"simple_neighbors", "attached_neighbors", "not_neighbors"};
    std::string js = "{";
    char buf[512];
    snprintf(buf, sizeof(buf), "\"gpus\": %d, \"level\": %d, \"error_control\": \"%s\", \"all_classes_ms\": %.6f, \"classes\": [", lastGpus, lastLevel,
             lastLevel < 0 ? "automatic" : "fixed", lastGpus > 1 ? allClassesMs : summary[0].ms + summary[1].ms + summary[2].ms);
    js += buf;
    for (int k = 0; k < 3; ++k) {
        const ClassSummary &c = summary[k];
        snprintf(buf, sizeof(buf), "%s{\"class\": \"%s\", \"tasks\": %lld, \"ms\": %.6f, \"last_round\": %d, \"unconverged\": [%lld, %lld, %lld, %lld, %lld], "
                                   "\"delta_max\": %s, \"delta_mean\": %s}",
                 k ? ", " : "", names[k], c.tasks, c.ms, c.lastRound, c.unconverged[1], c.unconverged[2], c.unconverged[3], c.unconverged[4], c.unconverged[5],
                 c.deltaMax < 0 ? "null" : std::to_string(c.deltaMax).c_str(), c.deltaMean < 0 ? "null" : std::to_string(c.deltaMean).c_str());
        js += buf;
    }
    js += "]}";
    return js;
}
