#!/bin/bash
# Round 2, 8 GPUs: bench N=8 weak with the gated stopwatch
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --no-cpu-baseline --no-strong > gpurun_out/r2af_bench_n8.json 2> gpurun_out/r2af_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r2af_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2af_bench_n8.json').read()); e=d['e2e']
print('N8 weak value', round(d['value']/1e9,2), 'ms', round(d['ms_per_step'],4), 'spread', d['step_ms_spread'], 'e2e idx', round(e['value']/1e9,2), 'composed', round(e['composed_input']['value']/1e9,2), 'lattice', round(e['lattice_input']['value']/1e9,2), 'clocks', d['clocks'])"
