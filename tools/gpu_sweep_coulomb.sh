#!/bin/bash
# parity, then the two S sectors with the Coulomb factorisation off / always / automatic
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
run() { echo "== $SHAPE $*"; env "$@" timeout 600 python tools/prof_case.py $SHAPE 1 2>&1 | grep -E "shape|Error|error" | tail -1; }
SHAPE=1000,100,4000,1100
run AGF2_B200_COULOMB=0
run AGF2_B200_COULOMB=2
run AGF2_B200_COULOMB=2 AGF2_B200_K2_MASK=0
SHAPE=100,1000,4000,1100
run AGF2_B200_COULOMB=0
run AGF2_B200_COULOMB=1
run AGF2_B200_COULOMB=2
