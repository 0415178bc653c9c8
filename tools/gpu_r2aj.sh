#!/bin/bash
# Round 2, GPU round trip AJ (1 GPU): containment split from 640 Ki queries per launch: tests, bench both arms, ncu s1/s2/contain of the 1M step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2aj_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aj_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2aj_bench.err
timeout 200 python bench.py --impl reference > gpurun_out/r2aj_bench_ref.json 2> gpurun_out/r2aj_bench_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2aj_bench.json').read()); e=d['e2e']; r=json.loads(open('gpurun_out/r2aj_bench_ref.json').read())
print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'lattice', round(e['lattice_input']['value']/1e9,2), 'strong', (d.get('strong') or {}).get('ms_per_batch'), 'mism', d['cpu_baseline']['mismatches_vs_gpu'], 'same config', d['config']==r['config'], 'launches', d['gpu_launches'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s[12]_ --launch-skip 18 -c 3 -f -o gpurun_out/r2aj_s1_s2 python tools/size_sweep.py 1048576 > gpurun_out/r2aj_ncu_s1s2.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2aj_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r2aj_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
