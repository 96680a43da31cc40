#!/bin/bash
# parity + planner sweep (LPT order / balanced column blocks / wave-tuned chunk size / workspace size)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
export AGF2_B200_VERBOSE=1
run() { echo "== $*"; env "$@" python tools/prof_case.py $SHAPE 2 2>&1 | grep -E "shape|items per" | tail -3; }
for SHAPE in 200,100,4000,1100 100,1000,4000,1100; do
  for OV in 0 1; do
    run AGF2_B200_OVERLAP=$OV AGF2_B200_LPT=0 AGF2_B200_BALANCED=0 AGF2_B200_TUNE_CHUNK=0
    run AGF2_B200_OVERLAP=$OV AGF2_B200_LPT=1 AGF2_B200_BALANCED=0 AGF2_B200_TUNE_CHUNK=1
    run AGF2_B200_OVERLAP=$OV AGF2_B200_LPT=1 AGF2_B200_BALANCED=1 AGF2_B200_TUNE_CHUNK=1
    run AGF2_B200_OVERLAP=$OV AGF2_B200_LPT=0 AGF2_B200_BALANCED=1 AGF2_B200_TUNE_CHUNK=1
    run AGF2_B200_OVERLAP=$OV AGF2_B200_LPT=1 AGF2_B200_BALANCED=1 AGF2_B200_TUNE_CHUNK=1 AGF2_B200_WORKSPACE_MB=47
    run AGF2_B200_OVERLAP=$OV AGF2_B200_LPT=1 AGF2_B200_BALANCED=1 AGF2_B200_TUNE_CHUNK=1 AGF2_B200_WORKSPACE_MB=32
  done
done 2>&1 | tee gpurun_out/sweep8.log
