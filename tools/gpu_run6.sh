#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for ov in 0 1; do for ws in 24 40 64; do
echo "== overlap=$ov ws=$ws"
AGF2_B200_OVERLAP=$ov AGF2_B200_WORKSPACE_MB=$ws python tools/prof_case.py 64,100,4000,1100 2 | tail -1
AGF2_B200_OVERLAP=$ov AGF2_B200_WORKSPACE_MB=$ws python tools/prof_case.py 16,1000,4000,1100 2 | tail -1
done; done
