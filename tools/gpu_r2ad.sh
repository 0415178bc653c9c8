#!/bin/bash
# Round 2, GPU round trip AD (1 GPU): config 4 detect bench with burst 8 (libjagua_cuda.so) and 16 (alt2)
mkdir -p gpurun_out
for lib in libjagua_cuda.so libjagua_cuda_alt2.so; do
JG_LIB_NAME=$lib timeout 600 python tools/bench_detect.py --check 100000 > gpurun_out/r2ad_detect_config4_$lib.jsonl 2> gpurun_out/r2ad_detect_config4_$lib.err; echo "config4 rc=$?"
python -c "
import json
for l in open('gpurun_out/r2ad_detect_config4_$lib.jsonl'):
    d=json.loads(l); print('$lib', d['workload'], 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'path', d['path'], 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'])"
done
