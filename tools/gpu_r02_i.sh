#!/bin/bash
# Round-2 GPU check I (1 GPU): does the (xi|ja) workspace stay in L2 between k1_form and k2_moment?  DRAM bytes of
# consecutive launches without ncu's cache flush (--cache-control none, one pass), for the plain stores, for stores
# with the L2 evict_last hint, and with a persisting-L2 set-aside on top; and the un-profiled time of each variant.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
SHAPE=64,100,4000,1100
export AGF2_B200_COULOMB=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for V in "0 0" "1 0" "1 80"; do
  set -- $V
  export AGF2_B200_W_EVICT_LAST=$1 AGF2_B200_L2_PERSIST_MB=$2
  TAG=el$1_p$2
  python tools/prof_case.py $SHAPE 3 2>&1 | grep shape | tail -1 | sed "s/^/[$TAG] /" | tee -a gpurun_out/r02i_times.log
  python tools/prof_case.py 1000,100,4000,1100 1 2>&1 | grep shape | tail -1 | sed "s/^/[$TAG] /" | tee -a gpurun_out/r02i_times.log
  ncu --cache-control none --clock-control none --metrics $M -k regex:'k1_form|k2_moment' -s 4 -c 12 --csv \
      --log-file gpurun_out/r02i_l2_$TAG.csv python tools/prof_case.py $SHAPE 1 > gpurun_out/r02i_ncu_$TAG.log 2>&1
done
