#!/bin/bash
# Round 2, GPU round trip AG (1 GPU): config 4 with every query compared on the final build (bursts on long-run layouts)
mkdir -p gpurun_out
timeout 600 python tools/bench_detect.py > gpurun_out/r2ag_detect_config4.jsonl 2> gpurun_out/r2ag_detect_config4.err; echo "config4 rc=$?"
python -c "
import json
for l in open('gpurun_out/r2ag_detect_config4.jsonl'):
    d=json.loads(l); print(d['workload'], 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'path', d['path'], 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'], 'cpu', round(d['cpu_baseline']['value']/1e6,1), 'frac', (d.get('roofline') or {}).get('frac'))"
