#!/bin/bash
# Round 2, GPU round trip I (1 GPU): one-batch-at-a-time latency vs CTAs per SM of the chunk kernels (co-resident chunks)
mkdir -p gpurun_out
for c in 0 2 3; do
  JG_CTAS_PER_SM=$c timeout 300 python bench.py --no-cpu-baseline --no-strong > gpurun_out/r2i_bench_ctas$c.json 2> gpurun_out/r2i_bench_ctas$c.err
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2i_bench_ctas$c.json') if l.startswith('{')][0]; e=d['e2e']
print('ctas/SM $c: value', round(d['value']/1e9,2), 'e2e idx', round(e['value']/1e9,2), 'serial ms idx', round(e['one_batch_at_a_time']['ms_per_step'],3), 'composed', round(e['composed_input']['one_batch_at_a_time']['ms_per_step'],3), 'lattice', round(e['lattice_input']['one_batch_at_a_time']['ms_per_step'],3))"
done
