// Evaluator3D — drop-in for the reference's abstract evaluator (/root/reference/src/evaluators/evaluator3d.cuh:30-260):
// builds the ordered task lists, owns the result buffers, drives the three per-class virtuals (the plugin API),
// computes the (i,j)/(j,i) defect and writes the csv / plain-text exports in the reference's format.
#ifndef EVALUATOR3D_CUH
#define EVALUATOR3D_CUH

#include <string>
#include <vector>

#include "../Mesh3d.cuh"
#include "../NumericalIntegrator3d.cuh"
#include "../common/gpu_timer.cuh"

// binary (not in the reference): full-precision records keyed (i, j), see outputResultsToFile
enum class output_format_enum { plainText = 1, csv = 2, binary = 3 };

class Evaluator3D {
public:
    Evaluator3D(const Mesh3D &mesh_, NumericalIntegrator3D &numIntegrator_);
    virtual ~Evaluator3D() = default;

    virtual void integrateOverSimpleNeighbors() = 0;
    virtual void integrateOverAttachedNeighbors() = 0;
    virtual void integrateOverNotNeighbors() = 0;

    virtual void runAllPairs(bool checkCorrectness = false);
    void runPairs(const std::vector<int3> &userSimpleNeighborsTasks, const std::vector<int3> &userAttachedNeighborsTasks,
                  const std::vector<int3> &userNotNeighborsTasks);
    bool outputResultsToFile(neighbour_type_enum neighborType, output_format_enum outputFormat) const;

    const deviceVector<int3> *getTasks(neighbour_type_enum t) const {
        switch (t) {
        case neighbour_type_enum::simple_neighbors: return &simpleNeighborsTasks;
        case neighbour_type_enum::attached_neighbors: return &attachedNeighborsTasks;
        case neighbour_type_enum::not_neighbors: return &notNeighborsTasks;
        default: return nullptr;
        }
    }
    // machine-readable summary of the last run (classes, counts, times, rounds, (i,j)/(j,i) defect summary): what the CLI
    // writes when env I2_SUMMARY_JSON names a file
    std::string getRunSummaryJson() const;
    // device views used by parity harnesses (the reference keeps these members protected)
    const deviceVector<Point3> *getResultsVector(neighbour_type_enum t) const;
    const deviceVector<double4> *getIntegralsVector(neighbour_type_enum t) const;
    const deviceVector<double> *getErrorsVector(neighbour_type_enum t) const;

protected:
    // Runge comparison is fused into i2_integrate_class; kept so that subclasses written against the reference compile.
    int compareIntegrationResults(neighbour_type_enum neighborType, bool allPairs = false);

    deviceVector<int3> simpleNeighborsTasks, attachedNeighborsTasks, notNeighborsTasks;
    deviceVector<double4> d_simpleNeighborsIntegrals, d_attachedNeighborsIntegrals, d_notNeighborsIntegrals;
    deviceVector<Point3> d_simpleNeighborsResults, d_attachedNeighborsResults, d_notNeighborsResults;

    const Mesh3D &mesh;
    NumericalIntegrator3D &numIntegrator;
    GpuTimer timer;
    // true while the task vectors hold the ordered lists of runAllPairs ([pairs ; reversed pairs]): the per-class virtuals then
    // use i2_integrate_pairs, whose results do not depend on how many GPUs a run uses
    bool tasksArePairs = false;
    // per-class record of the last run, filled by the per-class virtuals / the multi-GPU driver
    struct ClassSummary {
        long long tasks = 0;
        double ms = 0.0;
        int lastRound = 0;
        long long unconverged[6] = {0, 0, 0, 0, 0, 0};
        double deltaMax = -1.0, deltaMean = -1.0;   // -1: not computed (no --checkresults)
    } summary[3];

private:
    void allocateClass(int cls, int taskCount);
    void runAllPairsMultiGpu(bool checkCorrectness);
    void reportDefects(int cls, double maxDelta, double meanDelta, long long n);
    deviceVector<double> simpleNeighborsErrors, attachedNeighborsErrors, notNeighborsErrors;
    bool distributed = false;     // the last run left its per-pair results row-striped over the GPUs (env I2_GPUS > 1)
    bool distributedChecked = false;
    long long distributedCount[3] = {0, 0, 0};
    int lastGpus = 1, lastLevel = 0;
    double allClassesMs = 0.0;
};

#endif  // EVALUATOR3D_CUH
