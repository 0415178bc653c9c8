#!/bin/bash
# One GPU round trip: parity tests, launch list, bench line. Usage: tools/gpu_check.sh  (run under gpurun)
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_nocpu.json')); print('ms/step', d['ms_per_step'], 'value', d['value']/1e9, 'e2e', d['e2e']['value']/1e9)"
