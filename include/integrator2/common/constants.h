// Compile-time constants of the integrator2 API.  The VALUES are part of the numerics (they select
// branches and the Runge tolerance) and are therefore identical to the reference's
// (/root/reference/src/common/constants.h:11-65); the kernels use the same values from csrc/i2_vec.cuh.
#ifndef CONSTANTS_H
#define CONSTANTS_H

struct CONSTANTS {
    // tolerances
    static constexpr double DOUBLE_MIN = 2e-6;       // zero test inside sign()/arg()/divide()
    static constexpr double EPS_ZERO = 1e-6;         // zero test of angles / lengths in the integration formulas
    static constexpr double EPS_ZERO2 = 1e-10;       // EPS_ZERO squared (rounded as in the reference)
    static constexpr double EPS_PSI_THETA = EPS_ZERO;
    static constexpr double EPS_PSI_THETA2 = EPS_PSI_THETA * EPS_PSI_THETA;
    static constexpr double EPS_INTEGRATION = 1e-5;  // Runge rule tolerance
    // math
    static constexpr double ONE_THIRD = 0.3333333333333333;
    static constexpr double PI = 3.14159265358979323846;
    static constexpr double TWO_PI = 6.28318530717958647692;
    static constexpr double RECIPROCAL_FOUR_PI = 0.079577471545947667884;
    // capacities / limits
    static constexpr int MAX_SIMPLE_NEIGHBORS_PER_CELL = 12;         // unused here: lists are sized exactly
    static constexpr int MAX_AUTO_REFINEMENT_TASK_COEFFICIENT = 4;   // unused here: refined tasks are never materialised
    static constexpr double MEMORY_REALLOCATION_COEFFICIENT = 1.25;
    static constexpr int MAX_REFINE_LEVEL = 5;
    static constexpr int MAX_GAUSS_POINTS = 13;
};

#endif  // CONSTANTS_H
