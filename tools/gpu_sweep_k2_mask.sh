#!/bin/bash
# parity with the generalised k2_moment + Coulomb factorisation, then timing with the factorisation on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
export AGF2_B200_VERBOSE=1
run() { echo "== $*"; env "$@" timeout 300 python tools/prof_case.py $SHAPE 2 2>&1 | grep -E "shape|Error|error" | tail -1; }
for SHAPE in 200,100,4000,1100 100,1000,4000,1100; do
  for OV in 0 1; do
    run AGF2_B200_OVERLAP=$OV AGF2_B200_COULOMB=0 AGF2_B200_K2_MASK=0
    run AGF2_B200_OVERLAP=$OV AGF2_B200_COULOMB=0 AGF2_B200_K2_MASK=1
    run AGF2_B200_OVERLAP=$OV AGF2_B200_COULOMB=1 AGF2_B200_K2_MASK=0
    run AGF2_B200_OVERLAP=$OV AGF2_B200_COULOMB=1 AGF2_B200_K2_MASK=1
  done
done 2>&1 | tee gpurun_out/sweep9.log
