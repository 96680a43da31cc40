#!/bin/bash
mkdir -p gpurun_out
SHAPE=${1:-64,100,4000,1100}
python tools/prof_case.py $SHAPE 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/prof_plain.log
python tools/prof_case.py $SHAPE 1 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_form -s 2 -c 2 -o gpurun_out/prof_k1 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k1.log 2>&1
python tools/prof_case.py $SHAPE 1 > gpurun_out/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k2_moment -s 2 -c 2 -o gpurun_out/prof_k2 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/ncu_k1.log gpurun_out/ncu_k2.log
ls -la gpurun_out/
