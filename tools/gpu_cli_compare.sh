#!/bin/bash
# The drop-in CLI against the reference's CLI on the same meshes (their own "Time for ... integration" lines)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
EX=$PWD/oracle/_ref/examples
OURS=$PWD/integrator2_b200/host/integrator2test3D
REF=$PWD/oracle/_ref/integrator2test3D
cd /tmp
for args in "-f $EX/Vint16k.dat -r 0" "-f $EX/s5m2.dat -s 0.0005" "-f $EX/s5m.dat -s 0.0005 -r 1 -c"; do
  echo "### $args"; echo "--- ours"; timeout 600 $OURS $args | grep -E "Time for|Out of|Iteration [0-5]," | tr '\n' ';'; echo
  echo "--- reference"; timeout 900 $REF $args | grep -E "Time for|Out of|Iteration [0-5]," | tr '\n' ';'; echo
done
echo "### ours only: Vint16k adaptive (the reference regenerates tasks in O(T*C) and is not run here)"
timeout 600 $OURS -f $EX/Vint16k.dat | grep -E "Time for|Out of|Iteration [0-5]," | tr '\n' ';'; echo
