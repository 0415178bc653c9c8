#!/bin/bash
# Round 2, GPU round trip T (1 GPU): per-batch piece schedule of the host-buffer path vs fixed pieces; tests; bench
mkdir -p gpurun_out
timeout 600 python tools/schedule_probe.py 2>&1 | tee gpurun_out/r2t_schedule_probe.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read()); e=d['e2e']
print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'composed serial', round(e['composed_input']['one_batch_at_a_time']['ms_per_step'],3), 'lattice serial', round(e['lattice_input']['one_batch_at_a_time']['ms_per_step'],3), 'strong', d.get('strong',{}).get('ms_per_batch'), 'mism', d['cpu_baseline']['mismatches_vs_gpu'])"
