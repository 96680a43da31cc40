#!/bin/bash
# Round-2 GPU check H (8 GPUs): the L shape (BASELINE configs[4]: nocc=250 nvir=2500 naux=10000 nphys=2750, SCS os=6/5
# ss=1/3) pole-sharded over the 8 GPUs: occupied sector complete, virtual sector sampled (2 visiting sub-blocks per ring
# step) and extrapolated by the pair count; projection parity of both sectors.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
S1=$(date +%s)
timeout 1100 $TR --master-port 29562 bench.py --gpus 8 --shape 250,2500,10000,2750 --os-factor 1.2 --ss-factor 0.3333333333333333 \
    --sharded --sample-limit 2 --steps 1 --warmup 2 > gpurun_out/r02h_bench_L_8gpu.json 2> gpurun_out/r02h_bench_L_8gpu.err
echo "L 8gpu rc=$? wall=$(( $(date +%s) - S1 )) s" | tee gpurun_out/r02h_wall.txt
nvidia-smi --query-gpu=index,memory.used --format=csv >> gpurun_out/r02h_wall.txt
tail -c 3500 gpurun_out/r02h_bench_L_8gpu.json; grep -v "OpenBLAS\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02h_bench_L_8gpu.err | tail -8
