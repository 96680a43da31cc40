#!/bin/bash
# Round-2 GPU check G (2 GPUs): bench --sharded on a reduced shape, complete and sampled (validates the sector-sequential
# sharded arm before the 8-GPU L run).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for L in 0 2; do
timeout 600 $TR --master-port 2957$L bench.py --gpus 2 --shape 60,600,2000,660 --os-factor 1.2 --ss-factor 0.3333333333333333 --sharded --sample-limit $L --steps 1 --warmup 2 \
   > gpurun_out/r02g_bench_small_sharded_$L.json 2> gpurun_out/r02g_bench_small_sharded_$L.err
echo "rc=$?"; tail -c 2500 gpurun_out/r02g_bench_small_sharded_$L.json; grep -v "OpenBLAS\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02g_bench_small_sharded_$L.err | tail -6
done
