#!/bin/bash
# Round 2, GPU round trip P (1 GPU): does the image layout (vs the packed blob) cost kernel time? bench A/B, then tests
mkdir -p gpurun_out
for v in 0 1 0 1; do
  JG_INCREMENTAL=$v timeout 300 python bench.py --no-cpu-baseline --no-strong > gpurun_out/r2p_bench_inc$v.json 2> gpurun_out/r2p_bench_inc$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2p_bench_inc$v.json').read()); e=d['e2e']
print('JG_INCREMENTAL=$v value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3))"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest_gpu.log
for v in 1; do
  JG_INCREMENTAL=$v timeout 300 python tools/update_probe.py 2>&1 | grep -v "collect (surrogate" > gpurun_out/r2p_update_probe_inc$v.txt; cat gpurun_out/r2p_update_probe_inc$v.txt
done
