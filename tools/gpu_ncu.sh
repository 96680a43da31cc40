#!/bin/bash
# ncu: launch list of a reduced-shape bench command + full captures of k1_form and the three k2_moment uses
mkdir -p gpurun_out
BSHAPE=30,300,1000,330
python bench.py --shape $BSHAPE --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_reduced_plain.json 2> gpurun_out/bench_reduced_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_r01c.csv python bench.py --shape $BSHAPE --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1
tail -c 1500 gpurun_out/bench_reduced_plain.json
SHAPE=64,100,4000,1100
export AGF2_B200_COULOMB=1
python tools/prof_case.py $SHAPE 1 > gpurun_out/prof_plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_form -s 4 -c 2 -o gpurun_out/prof_k1_r01c -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_moment -s 61 -c 6 -o gpurun_out/prof_k2_r01c -f python tools/prof_case.py $SHAPE 1 > gpurun_out/ncu_k2.log 2>&1
cat gpurun_out/prof_plain_c.log | tail -2; tail -2 gpurun_out/ncu_k1.log gpurun_out/ncu_k2.log
unset AGF2_B200_COULOMB
for S in 285,507,700,264 507,285,700,264; do
  for T in 0 1; do echo "== $S THIN=$T"; AGF2_B200_THIN=$T python tools/prof_case.py $S 2 2>&1 | grep -E "shape|rror" | tail -1; done
done 2>&1 | tee gpurun_out/sweep_benzene.log
