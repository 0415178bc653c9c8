#!/bin/bash
# Round 2, 2 GPUs: bench N=2 weak with the final build (split containment from 640 Ki)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-cpu-baseline --no-strong > gpurun_out/r2am_bench_n2.json 2> gpurun_out/r2am_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2am_bench_n2.json').read()); e=d['e2e']
print('N=2 value', round(d['value']/1e9,2), 'ms', round(d['ms_per_step'],4), 'e2e idx', round(e['value']/1e9,2), 'lattice', round(e['lattice_input']['value']/1e9,2), 'launches', d['gpu_launches'])"
