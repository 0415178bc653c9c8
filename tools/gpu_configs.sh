#!/bin/bash
# BASELINE.json configs 1, 3, 4, 5 on one GPU, one JSON line each under gpurun_out/<tag>_*.jsonl.
# Usage (under gpurun): bash tools/gpu_configs.sh r1j
tag=${1:-run}
mkdir -p gpurun_out
timeout 400 python tools/bench_detect.py > gpurun_out/${tag}_config4.jsonl 2> gpurun_out/c4.err; echo "config 4 rc=$?"
timeout 500 python tools/bench_collect.py > gpurun_out/${tag}_collect_n1.jsonl 2> gpurun_out/col.err; echo "config 3 rc=$?"
rm -f gpurun_out/${tag}_config5_lookahead.jsonl
for i in jakobs1 jakobs2 albano; do
    timeout 300 python tools/bench_mspp.py --instance $i --lookahead 1 >> gpurun_out/${tag}_config5_lookahead.jsonl 2>> gpurun_out/c5.err
done; echo "config 5 rc=$?"
timeout 200 python tools/run_lbf.py --samples 500000 --solution gpurun_out/${tag}_swim_solution.json > gpurun_out/${tag}_lbf_config1.jsonl 2> gpurun_out/lbf.err; echo "config 1 rc=$?"
python - <<PY
import json
for f in ["${tag}_config4.jsonl", "${tag}_collect_n1.jsonl", "${tag}_config5_lookahead.jsonl", "${tag}_lbf_config1.jsonl"]:
    for l in open("gpurun_out/" + f):
        d = json.loads(l)
        print(f, {k: (round(v / 1e9, 3) if isinstance(v, (int, float)) and not isinstance(v, bool) and v > 1e8 else v) for k, v in d.items() if not isinstance(v, (dict, list))})
PY
