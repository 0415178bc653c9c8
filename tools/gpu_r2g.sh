#!/bin/bash
# Round 2, GPU round trip G (1 GPU): all parity tests; collect timing; config 3 (10M queries, every query compared)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest_gpu.log
timeout 120 python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2g_collect_shirts.txt 2>&1; tail -1 gpurun_out/r2g_collect_shirts.txt
timeout 120 python tools/collect_probe.py trousers_full 1048576 > gpurun_out/r2g_collect_trousers.txt 2>&1; tail -1 gpurun_out/r2g_collect_trousers.txt
timeout 120 python tools/collect_probe.py swim_half 262144 > gpurun_out/r2g_collect_swim.txt 2>&1; tail -1 gpurun_out/r2g_collect_swim.txt
timeout 900 python tools/bench_collect.py > gpurun_out/r2g_collect_config3_n1.jsonl 2> gpurun_out/r2g_collect_config3.err; echo "config3 rc=$?"; cut -c1-420 gpurun_out/r2g_collect_config3_n1.jsonl; tail -3 gpurun_out/r2g_collect_config3.err
python -c "
import json
for ln in open('gpurun_out/r2g_collect_config3_n1.jsonl'):
    d=json.loads(ln); print(d['workload'], 'G q/s', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value']/1e6, 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'])
"
