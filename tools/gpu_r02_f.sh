#!/bin/bash
# Round-2 GPU check F (1 GPU): the GPU test-suite; DRAM traffic of consecutive k1_form -> k2_moment launches WITHOUT the
# L2 flush ncu applies by default (--cache-control none, one pass, no kernel replay): is the (xi|ja) workspace served
# from L2?; the secondary nphys = 367 shape with the two-stream overlap off (k1_form timed alone).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02f_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02f_pytest_gpu.log
SHAPE=64,100,4000,1100
export AGF2_B200_COULOMB=1
python tools/prof_case.py $SHAPE 1 > gpurun_out/r02f_prof_plain.log 2>&1 && \
ncu --cache-control none --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum \
    -k regex:'k1_form|k2_moment' -s 4 -c 24 --csv --log-file gpurun_out/r02f_l2_k1_k2_nocachectl.csv \
    python tools/prof_case.py $SHAPE 1 > gpurun_out/r02f_ncu_l2.log 2>&1
tail -2 gpurun_out/r02f_prof_plain.log; tail -3 gpurun_out/r02f_ncu_l2.log
unset AGF2_B200_COULOMB
for S in 100,1000,4000,367 1000,100,4000,367; do
  for OV in 1 0; do echo "== $S OVERLAP=$OV"; AGF2_B200_OVERLAP=$OV python tools/prof_case.py $S 2 2>&1 | grep -E "shape|rror" | tail -1; done
done 2>&1 | tee gpurun_out/r02f_secondary_p367.log
