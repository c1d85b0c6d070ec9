"""TEST INFRASTRUCTURE: CPU oracle of the integrator2 hot path (see oracle/oracle.cpp)."""
