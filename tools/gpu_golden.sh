#!/bin/bash
# Runs the UNMODIFIED reference build (oracle/_ref) on the GPU box and writes full-precision binary dumps of its
# results to gpurun_out/golden/ (converted to tests/golden/*.npz by tools/pack_golden.py).  Usage: tools/gpu_golden.sh
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/golden
mkdir -p $OUT
EX=oracle/_ref/examples
R=oracle/_ref/ref_dump
RN=oracle/_ref/ref_dump_nofma
run() { # name mesh scale mode(-1|N) extra...
    local name=$1 mesh=$2 scale=$3 mode=$4; shift 4
    local rflag=""; [ "$mode" != "-1" ] && rflag="-r $mode"
    timeout 600 $R -f $EX/$mesh -s $scale $rflag -o $OUT/$name "$@" > $OUT/$name.log 2>&1 || echo "FAILED $name" >> $OUT/failures.txt
}
run G1_r0 G1.dat 1.0 0
run G1_r1 G1.dat 1.0 1
run G1_r2 G1.dat 1.0 2
run G1_ad G1.dat 1.0 -1
timeout 600 $RN -f $EX/G1.dat -r 0 -o $OUT/G1_r0_nofma > $OUT/G1_r0_nofma.log 2>&1
for f in $EX/Case*.dat $EX/G1Sosed.dat $EX/G1new.dat $EX/G1Cont.dat $EX/G1contact.dat $EX/G1contactR.dat $EX/genCase.dat $EX/Test.dat; do
    b=$(basename $f .dat)
    run ${b}_r0 $b.dat 1.0 0
    run ${b}_ad $b.dat 1.0 -1
done
run s5m_r0 s5m.dat 0.0005 0 --not-stride 101
run s5m_r1 s5m.dat 0.0005 1 --not-stride 101
run s5m_ad s5m.dat 0.0005 -1 --not-stride 101
timeout 600 $RN -f $EX/s5m.dat -s 0.0005 -r 0 -o $OUT/s5m_r0_nofma --not-stride 101 > $OUT/s5m_r0_nofma.log 2>&1
run cubehole_r0 cubehole.dat 1.0 0 --not-stride 7
run cubehole_ad cubehole.dat 1.0 -1 --not-stride 7
run ellipsoid2000_r0 ellipsoid2000 1.0 0 --not-stride 211
run extrafine_r0 1x1x1_extrafine 1.0 0 --not-stride 5003 --small-stride 5
run s5m2_ad s5m2.dat 0.0005 -1 --not-stride 1009 --small-stride 3
run Vint16k_r0 Vint16k.dat 1.0 0 --not-stride 20011 --small-stride 7
# the reference's own timing lines on the headline config (its CLI, not the dump harness)
( cd /tmp && timeout 900 $OLDPWD/oracle/_ref/integrator2test3D -f $OLDPWD/$EX/Vint16k.dat -r 0 ) > $OUT/ref_cli_Vint16k_r0.log 2>&1
( cd /tmp && timeout 900 $OLDPWD/oracle/_ref/integrator2test3D -f $OLDPWD/$EX/s5m2.dat -s 0.0005 ) > $OUT/ref_cli_s5m2_ad.log 2>&1
ls -la $OUT | tail -5
du -sh $OUT
