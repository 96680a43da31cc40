#!/bin/bash
# N-GPU: reduced-shape sanity of the gathered e2e path, then the full S bench
mkdir -p gpurun_out
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $NG --shape 25,241,1001,265 --steps 1 --warmup 1 > gpurun_out/bench_small_${NG}gpu.json 2> gpurun_out/bench_small_${NG}gpu.err || { echo small failed; tail -20 gpurun_out/bench_small_${NG}gpu.err; exit 1; }
grep -o '"e2e": {[^}]*}' gpurun_out/bench_small_${NG}gpu.json
timeout 1500 $TR bench.py --gpus $NG > gpurun_out/bench_S_${NG}gpu.json 2> gpurun_out/bench_S_${NG}gpu.err; echo rc=$?
grep -o '"value": [0-9.]*' gpurun_out/bench_S_${NG}gpu.json | head -1; grep -o '"e2e": {[^}]*}' gpurun_out/bench_S_${NG}gpu.json; tail -n 3 gpurun_out/bench_S_${NG}gpu.err
