#!/bin/bash
# N-GPU: sharded device path + sharded host-pointer e2e (full S bench, the driver's command line)
mkdir -p gpurun_out
NG=${NG:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG > gpurun_out/bench_S_${NG}gpu.json 2> gpurun_out/bench_S_${NG}gpu.err; echo rc=$?
tail -c 3500 gpurun_out/bench_S_${NG}gpu.json; tail -5 gpurun_out/bench_S_${NG}gpu.err
