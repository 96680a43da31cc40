#!/bin/bash
# first GPU contact: FP64 peaks, parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 auxgf_b200/lib/fp64_peak > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err; echo "fp64_peak rc=$?"
cat gpurun_out/fp64_peak.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
