#!/bin/bash
# final sanity of the tree: parity suite, smoke, whole-driver timing (compute-sanitizer is closed on this pool)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 420 python tools/driver_timing.py 264,21,700 3 > gpurun_out/driver_timing.log 2>&1; tail -n 6 gpurun_out/driver_timing.log
