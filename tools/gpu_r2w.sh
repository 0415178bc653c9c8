#!/bin/bash
# Round 2, GPU round trip W (1 GPU): k_collect_grid with lane = (query, terminal, entry) (libjagua_cuda.so) vs lane = (query, terminal) (alt, -DCG_TRIPLES=0); parity; bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest_gpu.log
for lib in libjagua_cuda.so libjagua_cuda_alt.so libjagua_cuda.so libjagua_cuda_alt.so; do
  JG_LIB_NAME=$lib timeout 300 python tools/bench_collect.py --queries 4000000 --reps 3 --no-check > gpurun_out/r2w_collect_$lib.jsonl 2> gpurun_out/r2w_collect_$lib.err
  python -c "
import json
for l in open('gpurun_out/r2w_collect_$lib.jsonl'):
    d=json.loads(l); print('$lib', d.get('workload'), 'value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,3))"
done
timeout 600 python tools/bench_collect.py > gpurun_out/r2w_collect_config3.jsonl 2> gpurun_out/r2w_collect_config3.err; echo "config3 rc=$?"
python -c "
import json
for l in open('gpurun_out/r2w_collect_config3.jsonl'):
    d=json.loads(l); print(d.get('workload'), 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'], 'frac', d.get('roofline',{}).get('frac'))"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read()); e=d['e2e']
print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'composed', round(e['composed_input']['value']/1e9,2), round(e['composed_input']['one_batch_at_a_time']['ms_per_step'],3), 'lattice', round(e['lattice_input']['value']/1e9,2), round(e['lattice_input']['one_batch_at_a_time']['ms_per_step'],3), 'strong', (d.get('strong') or {}).get('ms_per_batch'))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect --launch-skip 4 -c 2 -f -o gpurun_out/r2w_collect python tools/bench_collect.py --workloads shirts_full --queries 2000000 --reps 1 --no-check > gpurun_out/r2w_ncu_collect.log 2>&1; echo "ncu collect rc=$?"
