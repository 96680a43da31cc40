#!/bin/bash
# Round-2 GPU check M (1 GPU): quad tiles (11..14 row blocks) -- parity tests, nphys=367 sectors with the mapping on/off;
# `--set full` capture of the k2_moment uses with the 384 MiB chunks.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02m_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02m_pytest_gpu.log
for S in 100,1000,4000,367 1000,100,4000,367; do
  for TH in 1 0; do for OV in 1 0; do echo "== $S THIN=$TH OVERLAP=$OV"; AGF2_B200_THIN=$TH AGF2_B200_OVERLAP=$OV python tools/prof_case.py $S 2 2>&1 | grep -E "shape|rror" | tail -1; done; done
done 2>&1 | tee gpurun_out/r02m_quad_p367.log
SHAPE=64,100,4000,1100
export AGF2_B200_COULOMB=1
ncu --set full --clock-control none --import-source on -k regex:'k2_moment|k2_fixup' -s 2 -c 12 -o gpurun_out/r02m_prof_k2 -f python tools/prof_case.py $SHAPE 1 > gpurun_out/r02m_ncu_k2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02m_prof_k2.ncu-rep gpurun_out/r02_k2_moment_ncu.csv
tail -2 gpurun_out/r02m_ncu_k2.log
