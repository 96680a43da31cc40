#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --shape 40,400,1600,440 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_medium.json 2> gpurun_out/bench_medium.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_medium.json').read().strip().splitlines()[-1]); print('medium value',d['value'],'e2e',d['e2e'])
PY
timeout 1500 python bench.py > gpurun_out/bench_S_full2.json 2> gpurun_out/bench_S_full2.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_S_full2.json').read().strip().splitlines()[-1]); print('S value',d['value'],'tflops',d['tflops'],'e2e',d['e2e'],'cpu',d['cpu_baseline'])
PY
tail -3 gpurun_out/bench_S_full2.err
