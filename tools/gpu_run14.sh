#!/bin/bash
# k2_moment captures (pair moments, M build, Coulomb step), launch list of a factorised reduced-shape bench, then the
# full S bench and the reference arm (the driver's two commands)
mkdir -p gpurun_out
SHAPE=64,100,4000,1100
AGF2_B200_COULOMB=1 python tools/prof_case.py $SHAPE 1 > gpurun_out/prof_plain_c.log 2>&1 && \
AGF2_B200_COULOMB=1 ncu --set full --clock-control none --import-source on -k regex:k2_moment -s 50 -c 4 -o gpurun_out/prof_k2_r01d -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k2.log 2>&1
tail -n 2 gpurun_out/ncu_k2.log
BSHAPE=40,400,400,220
python bench.py --shape $BSHAPE --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_reduced_fact.json 2> gpurun_out/bench_reduced_fact.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_r01d.csv python bench.py --shape $BSHAPE --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1
tail -c 900 gpurun_out/bench_reduced_fact.json
timeout 2400 python bench.py > gpurun_out/bench_S_1gpu.json 2> gpurun_out/bench_S_1gpu.err; echo rc=$?
tail -c 3000 gpurun_out/bench_S_1gpu.json; tail -n 3 gpurun_out/bench_S_1gpu.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_S_reference.json 2> gpurun_out/bench_S_reference.err; echo rc=$?
cat gpurun_out/bench_S_reference.json
