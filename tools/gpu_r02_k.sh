#!/bin/bash
# Round-2 GPU check K (1 GPU): workspace (chunk) size sweep -- the L2-resident chunk was a design assumption that the
# ncu capture without cache flush does not support, so larger chunks (fewer, longer launches) may simply be better.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export AGF2_B200_VERBOSE=1
for SHAPE in 1000,100,4000,1100 100,1000,4000,1100; do
  for WS in 40 96 200 400; do
    echo "== $SHAPE WORKSPACE_MB=$WS"
    AGF2_B200_WORKSPACE_MB=$WS python tools/prof_case.py $SHAPE 1 2>&1 | grep -E "shape|items per|rror" | tail -3
  done
done 2>&1 | tee gpurun_out/r02k_workspace_sweep.log
echo "== overlap off"
for WS in 40 200; do AGF2_B200_OVERLAP=0 AGF2_B200_WORKSPACE_MB=$WS python tools/prof_case.py 200,100,4000,1100 2 2>&1 | grep -E "shape" | tail -1; done 2>&1 | tee -a gpurun_out/r02k_workspace_sweep.log
