#!/bin/bash
# Round 2, GPU round trip E (1 GPU): all parity tests, bench, collect timing, ncu of both collect kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2e_bench.json')); print('ms/step', d['ms_per_step'], 'value', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, 'serial', d['e2e']['one_batch_at_a_time']['ms_per_step'])"
timeout 120 python tools/size_sweep.py > gpurun_out/r2e_size_sweep.txt 2>&1; cat gpurun_out/r2e_size_sweep.txt
timeout 120 python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2e_collect_shirts.txt 2>&1; tail -9 gpurun_out/r2e_collect_shirts.txt
timeout 120 python tools/collect_probe.py trousers_full 1048576 > gpurun_out/r2e_collect_trousers.txt 2>&1; tail -1 gpurun_out/r2e_collect_trousers.txt
timeout 120 python tools/collect_probe.py swim_half 262144 > gpurun_out/r2e_collect_swim.txt 2>&1; tail -1 gpurun_out/r2e_collect_swim.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect --launch-skip 6 -c 2 -f -o gpurun_out/r2e_collect_shirts python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2e_ncu_collect.log 2>&1; echo "ncu rc=$?"
