#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/prof_case.py 64,100,4000,1100 2
python tools/prof_case.py 16,1000,4000,1100 2
