#!/bin/bash
# Round 2, GPU round trip AE (1 GPU): bench with the gated stopwatch; the timer tests
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2ae_bench.err
timeout 200 python bench.py --impl reference > gpurun_out/r2ae_bench_ref.json 2> gpurun_out/r2ae_bench_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2ae_bench.json').read()); e=d['e2e']; r=json.loads(open('gpurun_out/r2ae_bench_ref.json').read())
print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'spread', d['step_ms_spread'], 'frac', round(d['roofline']['frac'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'mism', d['cpu_baseline']['mismatches_vs_gpu'], 'same config', d['config']==r['config'], 'launches', d['gpu_launches'])"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ae_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
