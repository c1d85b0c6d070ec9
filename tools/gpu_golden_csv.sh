#!/bin/bash
# TEST INFRASTRUCTURE: the reference's own CLI (oracle/_ref/integrator2test3D, unmodified sources built by oracle/build_ref.sh)
# writes its csv exports for G1 `-r 0 --exporttocsv` on the GPU box; tools/pack_golden_csv.py packs them into
# tests/golden/G1_r0_csv.npz (rows as written: the test sorts by (i, j), the reference's row order is atomicAdd order).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/golden_csv && cd gpurun_out/golden_csv
"$GRAFT_REPO_ROOT/oracle/_ref/integrator2test3D" -f "$GRAFT_REPO_ROOT/oracle/_ref/examples/G1.dat" -r 0 --exporttocsv > ref_stdout.txt 2>&1
ls -la
