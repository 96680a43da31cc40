#!/bin/bash
# Round-2 GPU check L (1 GPU): whole-driver timing (benzene cc-pVTZ size model) and the refreshed ncu evidence:
# launch list of a reduced-shape bench command, `--set full` captures of k1_form and the k2_moment uses.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02l_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02l_pytest_gpu.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/r02l_bench_S_short.json 2> gpurun_out/r02l_bench_S_short.err
tail -c 1500 gpurun_out/r02l_bench_S_short.json; tail -3 gpurun_out/r02l_bench_S_short.err
python tools/driver_timing.py > gpurun_out/r02l_driver_timing.log 2>&1; tail -6 gpurun_out/r02l_driver_timing.log
BSHAPE=40,400,400,220
python bench.py --shape $BSHAPE --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02l_bench_reduced.json 2> gpurun_out/r02l_bench_reduced.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02l_launches_bench_40_400_400_220.csv python bench.py --shape $BSHAPE --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02l_ncu_launches.log 2>&1
tail -c 600 gpurun_out/r02l_bench_reduced.json
SHAPE=64,100,4000,1100
export AGF2_B200_COULOMB=1
python tools/prof_case.py $SHAPE 1 > gpurun_out/r02l_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_form -s 4 -c 2 -o gpurun_out/r02l_prof_k1 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/r02l_ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k2_moment|k2_fixup' -s 60 -c 10 -o gpurun_out/r02l_prof_k2 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/r02l_ncu_k2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02l_prof_k1.ncu-rep gpurun_out/r02_k1_form_ncu.csv
python tools/ncu_summary.py gpurun_out/r02l_prof_k2.ncu-rep gpurun_out/r02_k2_moment_ncu.csv
tail -2 gpurun_out/r02l_prof_plain.log gpurun_out/r02l_ncu_k1.log gpurun_out/r02l_ncu_k2.log
