// EvaluatorJ3DK host class: one C-ABI call per neighbour class + the reference's progress lines.
#include "../../include/integrator2/evaluators/evaluatorJ3DK.cuh"
#include "host_context.h"

void EvaluatorJ3DK::integrateClass(neighbour_type_enum neighborType) {
    const int cls = (int)neighborType;
    const deviceVector<int3> *tasks = getTasks(neighborType);
    const deviceVector<double4> *integrals = getIntegralsVector(neighborType);
    const deviceVector<Point3> *results = getResultsVector(neighborType);
    const bool adaptive = numIntegrator.getErrorControlType() == error_control_type_enum::automatic_error_control;
    const int level = adaptive ? I2_LEVEL_ADAPTIVE : numIntegrator.getFixedRefinementLevel();
    const int n = tasks->size;

    unsigned char *refinements = nullptr, *converged = nullptr;
    if (adaptive) {
        refinements = numIntegrator.getRefinementsRequired(neighborType)->data;
        converged = numIntegrator.getIntegralsConverged(neighborType)->data;
    }
    i2_stats st;
    if (tasksArePairs)   // the ordered list of runAllPairs: results independent of the number of GPUs a run uses
        checkI2Errors(i2_integrate_pairs(i2host::context(), cls, (const int *)tasks->data, n / 2, level, (double *)integrals->data,
                                         (double *)results->data, refinements, converged, &st));
    else
        checkI2Errors(i2_integrate_class(i2host::context(), cls, (const int *)tasks->data, n, level, (double *)integrals->data,
                                         (double *)results->data, refinements, converged, &st));
    summary[cls].lastRound = st.last_round;
    for (int m = 0; m < 6; ++m) summary[cls].unconverged[m] = st.unconverged[m];
    if (adaptive) {
        // the lines the reference prints while it iterates (src/evaluators/evaluatorJ3DK.cu:956,980; evaluator3d.cu:338)
        printf("Iteration 0, integrating %d tasks\n", n);
        long long checked = n;
        for (int m = 1; m <= st.last_round; ++m) {
            printf("Iteration %d, integrating %d tasks\n", m, (int)st.integrated[m]);
            printf("Out of %d tasks: %d converged, %d did not converge\n", (int)checked, (int)(checked - st.unconverged[m]),
                   (int)st.unconverged[m]);
            checked = st.unconverged[m];
        }
    }
}

void EvaluatorJ3DK::integrateOverSimpleNeighbors() {
    printf("\nIntegrating over simple neighbors (%d pairs)...\n", simpleNeighborsTasks.size);
    integrateClass(neighbour_type_enum::simple_neighbors);
}

void EvaluatorJ3DK::integrateOverAttachedNeighbors() {
    printf("\nIntegrating over attached neighbors (%d pairs)...\n", attachedNeighborsTasks.size);
    integrateClass(neighbour_type_enum::attached_neighbors);
}

void EvaluatorJ3DK::integrateOverNotNeighbors() {
    printf("\nIntegrating over not neighbors (%d pairs)...\n", notNeighborsTasks.size);
    integrateClass(neighbour_type_enum::not_neighbors);
}
