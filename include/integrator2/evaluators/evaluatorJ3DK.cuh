// EvaluatorJ3DK — integrals of the gradient of the Newtonian potential J_3D(K_i, K_j); drop-in for
// /root/reference/src/evaluators/evaluatorJ3DK.cuh:140-211.  Each integrateOver* is ONE call of the C ABI's
// i2_integrate_class (regular-part quadrature incl. refinement / error control, closed-form singular part,
// final assembly) followed by the same stdout lines the reference prints.
#ifndef EVALUATORJ3DK_CUH
#define EVALUATORJ3DK_CUH

#include "evaluator3d.cuh"

class EvaluatorJ3DK : public Evaluator3D {
public:
    EvaluatorJ3DK(const Mesh3D &mesh_, NumericalIntegrator3D &numIntegrator_) : Evaluator3D(mesh_, numIntegrator_) {}

    void integrateOverSimpleNeighbors() override;
    void integrateOverAttachedNeighbors() override;
    void integrateOverNotNeighbors() override;

private:
    void integrateClass(neighbour_type_enum neighborType);
};

#endif  // EVALUATORJ3DK_CUH
