#!/bin/bash
# bench.py under torchrun on N GPUs of one box (the driver's launch line), output to gpurun_out/r02_bench_${N}gpu.json
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -2 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_${N}gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['secondary']['strong_scaling_c5'])
print(d['secondary_c3']['strong_scaling_c3'])
PY
