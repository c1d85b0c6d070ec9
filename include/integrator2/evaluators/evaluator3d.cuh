// Evaluator3D — drop-in for the reference's abstract evaluator (/root/reference/src/evaluators/evaluator3d.cuh:30-260):
// builds the ordered task lists, owns the result buffers, drives the three per-class virtuals (the plugin API),
// computes the (i,j)/(j,i) defect and writes the csv / plain-text exports in the reference's format.
#ifndef EVALUATOR3D_CUH
#define EVALUATOR3D_CUH

#include <vector>

#include "../Mesh3d.cuh"
#include "../NumericalIntegrator3d.cuh"
#include "../common/gpu_timer.cuh"

enum class output_format_enum { plainText = 1, csv = 2 };

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

private:
    void allocateClass(int cls, int taskCount);
    deviceVector<double> simpleNeighborsErrors, attachedNeighborsErrors, notNeighborsErrors;
};

#endif  // EVALUATOR3D_CUH
