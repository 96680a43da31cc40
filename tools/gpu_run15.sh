#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python tools/driver_timing.py 264,21,700 3 2>&1 | tee gpurun_out/driver_timing.log | tail -8
