#!/bin/bash
# Round 2, 4 GPUs: bench N=2 and N=4 weak with the final build (gated stopwatch), pool strong leg included
mkdir -p gpurun_out
for n in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/r2ah_bench_n$n.json 2> gpurun_out/r2ah_bench_n$n.err; echo "bench n$n rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/r2ah_bench_n$n.json').read()); e=d['e2e']; s=d.get('strong') or {}
print('N=$n value', round(d['value']/1e9,2), 'ms', round(d['ms_per_step'],4), 'e2e idx', round(e['value']/1e9,2), 'composed', round(e['composed_input']['value']/1e9,2), 'lattice', round(e['lattice_input']['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'pool strong ms', s.get('ms_per_batch'), s.get('same_verdicts_as_single_engine'))"
done
