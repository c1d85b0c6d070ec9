// NumericalIntegrator3D — drop-in for the reference class (/root/reference/src/NumericalIntegrator3d.cuh:59-620).
// It owns the quadrature rule, the error-control mode and the per-class bookkeeping (refinement counters,
// converged flags).  In this implementation refined control panels are NOT materialised for the integration
// itself (the kernels rebuild every child triangle on the fly); the refined mesh arrays behind
// getRefinedVertices()/getRefinedCells()/getRefinedCellMeasures() are produced for fixed levels because the CLI
// exports them (tests/integrator3D/main.cu:155-169).
#ifndef NUMERICAL_INTEGRATOR_3D_CUH
#define NUMERICAL_INTEGRATOR_3D_CUH

#include "Mesh3d.cuh"
#include "QuadratureFormula3d.cuh"
#include "common/constants.h"

enum class error_control_type_enum {
    fixed_refinement_level = 0,   // every control panel split N times
    automatic_error_control = 1   // Runge rule, per-task depth
};

class NumericalIntegrator3D {
public:
    NumericalIntegrator3D(const Mesh3D &mesh_, const QuadratureFormula3D &qf_);
    virtual ~NumericalIntegrator3D() = default;

    void setFixedRefinementLevel(int refinementLevel = 0);
    void prepareTasksAndMesh(const deviceVector<int3> &simpleNeighborsTasks, const deviceVector<int3> &attachedNeighborsTasks,
                             const deviceVector<int3> &notNeighborsTasks);

    // Kept for source compatibility with user-written evaluators.  The stock EvaluatorJ3DK does not call them:
    // gathering, refinement and the selection of cells happen inside i2_integrate_class.
    void gatherResults(deviceVector<double4> &results, neighbour_type_enum neighborType) const;
    void refineMesh(neighbour_type_enum updateTasksNeighborType = neighbour_type_enum::undefined);
    void resetMesh();
    int determineCellsToBeRefined(const deviceVector<int> &restTasks, const deviceVector<int3> *tasks, neighbour_type_enum neighborType);

    int getGaussPointsNumber() const { return GaussPointsNum; }
    int getQuadratureFormulaOrder() const { return qf.order; }
    error_control_type_enum getErrorControlType() const { return errorControlType; }
    int getFixedRefinementLevel() const { return meshRefinementLevel; }

    const deviceVector<int3> *getRefinedTasks(neighbour_type_enum t) const { return valid(t) ? &refinedTasks[(int)t] : nullptr; }
    const deviceVector<double4> *getResults(neighbour_type_enum t) const { return valid(t) ? &refinedResults[(int)t] : nullptr; }
    const auto &getRefinedVertices() const { return refinedVertices; }
    const auto &getRefinedCells() const { return refinedCells; }
    const auto &getRefinedCellMeasures() const { return refinedCellMeasures; }
    auto &getCellsToBeRefined() const { return cellsToBeRefined; }
    deviceVector<unsigned char> *getIntegralsConverged(neighbour_type_enum t) { return valid(t) ? &integralsConverged[(int)t] : nullptr; }
    deviceVector<unsigned char> *getRefinementsRequired(neighbour_type_enum t) { return valid(t) ? &refinementsRequired[(int)t] : nullptr; }

private:
    static bool valid(neighbour_type_enum t) { return (int)t >= 0 && (int)t < 3; }

    const int GaussPointsNum;
    const Mesh3D &mesh;
    const QuadratureFormula3D &qf;
    error_control_type_enum errorControlType;
    int meshRefinementLevel = 0;

    deviceVector<Point3> refinedVertices;
    deviceVector<int3> refinedCells;
    deviceVector<double> refinedCellMeasures;
    deviceVector<int3> refinedTasks[3];            // left empty (never materialised)
    deviceVector<double4> refinedResults[3];       // left empty
    deviceVector<unsigned char> integralsConverged[3];
    deviceVector<unsigned char> refinementsRequired[3];
    deviceVector<int> cellsToBeRefined;
};

#endif  // NUMERICAL_INTEGRATOR_3D_CUH
