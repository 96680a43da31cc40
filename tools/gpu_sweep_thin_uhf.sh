#!/bin/bash
# parity incl. thin tiles, then timing: thin tiles on/off (S virtual sector, benzene-like self-consistent shape), UHF
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
run() { echo "== $SHAPE $*"; env "$@" timeout 600 python tools/prof_case.py $SHAPE $REPS 2>&1 | grep -E "shape|Error|error" | tail -1; }
REPS=2
for SHAPE in 285,507,700,264 507,285,700,264 200,100,4000,1100; do
  run AGF2_B200_THIN=0
  run AGF2_B200_THIN=1
done
REPS=1
SHAPE=1000,100,4000,1100
run AGF2_B200_THIN=1
echo "== UHF 100,1000,99,1001,4000,1100"
AGF2_B200_COULOMB=0 timeout 600 python tools/prof_case_uhf.py 2>&1 | tail -1
AGF2_B200_COULOMB=2 timeout 600 python tools/prof_case_uhf.py 2>&1 | tail -1
