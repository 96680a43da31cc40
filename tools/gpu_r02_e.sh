#!/bin/bash
# Round-2 GPU check E (8 GPUs): (1) the driver's command line at N = 8 on the S shape (fewer timed steps), wall time kept;
# (2) the L shape (BASELINE configs[4]: nocc=250 nvir=2500 naux=10000 nphys=2750, SCS os=6/5 ss=1/3) pole-sharded over
# the 8 GPUs: occupied sector complete, virtual sector sampled (2 visiting sub-blocks per ring step) and extrapolated;
# (3) the multi-rank parity tests at 8 ranks.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02e_gpu.txt 2>&1
free -g >> gpurun_out/r02e_gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
S0=$(date +%s)
timeout 870 $TR --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 5 > gpurun_out/r02e_bench_S_8gpu.json 2> gpurun_out/r02e_bench_S_8gpu.err
echo "S 8gpu rc=$? wall=$(( $(date +%s) - S0 )) s" | tee gpurun_out/r02e_wall.txt
tail -c 1800 gpurun_out/r02e_bench_S_8gpu.json; grep -v "OpenBLAS\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02e_bench_S_8gpu.err | tail -5
S1=$(date +%s)
timeout 900 $TR --master-port 29562 bench.py --gpus 8 --shape 250,2500,10000,2750 --os-factor 1.2 --ss-factor 0.3333333333333333 \
    --sharded --sample-limit 2 --steps 1 --warmup 2 > gpurun_out/r02e_bench_L_8gpu.json 2> gpurun_out/r02e_bench_L_8gpu.err
echo "L 8gpu rc=$? wall=$(( $(date +%s) - S1 )) s" | tee -a gpurun_out/r02e_wall.txt
tail -c 3000 gpurun_out/r02e_bench_L_8gpu.json; grep -v "OpenBLAS\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02e_bench_L_8gpu.err | tail -8
S2=$(date +%s)
(timeout 600 python -m pytest tests/test_multi_gpu.py -q -x -k "8") > gpurun_out/r02e_pytest_multi8.log 2>&1
echo "pytest multi 8 rc=$? wall=$(( $(date +%s) - S2 )) s" | tee -a gpurun_out/r02e_wall.txt
tail -4 gpurun_out/r02e_pytest_multi8.log
