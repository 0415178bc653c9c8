#!/bin/bash
# Round 2, GPU round trip N (1 GPU): incremental layout patches - all parity tests, update latency A/B (JG_INCREMENTAL=0/1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n_pytest_gpu.log
for v in 0 1; do
  echo "JG_INCREMENTAL=$v"; JG_INCREMENTAL=$v timeout 300 python tools/update_probe.py 2>&1 | grep -v "collect (surrogate" > gpurun_out/r2n_update_probe_inc$v.txt; cat gpurun_out/r2n_update_probe_inc$v.txt
done
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; wc -l gpurun_out/r2n_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read()); e=d['e2e']
print('value', round(d['value']/1e9,3), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'strong', d['strong']['one_batch_at_a_time'] if 'error' not in d['strong'] else d['strong'])"
