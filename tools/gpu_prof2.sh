#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
SHAPE=64,100,4000,1100
# launch list of a (reduced-shape) bench command: shares of k1_form / k2_moment in a step
python bench.py --shape 24,240,1000,264 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_small_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --shape 24,240,1000,264 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1
python tools/prof_case.py $SHAPE 1 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_form -s 2 -c 2 -o gpurun_out/prof_k1_v2 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k1.log 2>&1
python tools/prof_case.py $SHAPE 1 > gpurun_out/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k2_moment -s 2 -c 2 -o gpurun_out/prof_k2_v2 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k2.log 2>&1
tail -2 gpurun_out/ncu_k1.log gpurun_out/ncu_k2.log; cat gpurun_out/bench_small_plain.log | cut -c1-600
