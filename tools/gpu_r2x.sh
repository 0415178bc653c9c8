#!/bin/bash
# Round 2, GPU round trip X (1 GPU): k_collect_grid lane = (query, terminal, entry) with convergent test rounds (alt, -DCG_TRIPLES=2) vs plain loops (libjagua_cuda.so)
mkdir -p gpurun_out
JG_LIB_NAME=libjagua_cuda_alt.so timeout 600 python -m pytest tests -m gpu -x -q -k "collect or ref_case or seeded or smoke or update" > gpurun_out/r2x_pytest_gpu_alt.log 2>&1; echo "pytest(alt) rc=$?"; tail -3 gpurun_out/r2x_pytest_gpu_alt.log
for lib in libjagua_cuda.so libjagua_cuda_alt.so libjagua_cuda.so libjagua_cuda_alt.so; do
  JG_LIB_NAME=$lib timeout 300 python tools/bench_collect.py --queries 4000000 --reps 3 > gpurun_out/r2x_collect_$lib.jsonl 2> gpurun_out/r2x_collect_$lib.err
  python -c "
import json
for l in open('gpurun_out/r2x_collect_$lib.jsonl'):
    d=json.loads(l); print('$lib', d.get('workload'), 'value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,3), 'mism', d['mismatches_vs_oracle'])"
done
JG_LIB_NAME=libjagua_cuda_alt.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect_grid --launch-skip 2 -c 1 -f -o gpurun_out/r2x_collect_alt python tools/bench_collect.py --workloads shirts_full --queries 2000000 --reps 1 --no-check > gpurun_out/r2x_ncu_collect.log 2>&1; echo "ncu collect rc=$?"
