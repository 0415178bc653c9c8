#!/bin/bash
# Round 2, GPU final state (1 GPU): final-state validation - all parity tests, update latency, bench (+ reference arm), ncu launch list + s1/s2 summaries
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2fin_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2fin_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2fin_smoke.log
for v in 0 1; do
  echo "JG_INCREMENTAL=$v"; JG_INCREMENTAL=$v timeout 300 python tools/update_probe.py 2>&1 | grep -v "collect (surrogate" > gpurun_out/r2fin_update_probe_inc$v.txt; cat gpurun_out/r2fin_update_probe_inc$v.txt
done
timeout 300 python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "bench rc=$?"; wc -l gpurun_out/r2fin_bench.json
timeout 200 python bench.py --impl reference > gpurun_out/r2fin_bench_ref.json 2> gpurun_out/r2fin_bench_ref.err; echo "ref rc=$?"; wc -l gpurun_out/r2fin_bench_ref.json
python -c "
import json
d=json.loads(open('gpurun_out/r2fin_bench.json').read()); e=d['e2e']; r=json.loads(open('gpurun_out/r2fin_bench_ref.json').read())
print('value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3), 'cpu', round(d['cpu_baseline']['value']/1e6,1), 'mism', d['cpu_baseline']['mismatches_vs_gpu'], 'ref', round(r['value']/1e6,1), 'same config', d['config']==r['config'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r2fin_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s[12]_grid --launch-skip 12 -c 2 -f -o gpurun_out/r2fin_s1_s2 python tools/size_sweep.py 1048576 > gpurun_out/r2fin_ncu_s1s2.log 2>&1; echo "ncu s1s2 rc=$?"
timeout 600 python tools/bench_detect.py > gpurun_out/r2fin_detect_config4.jsonl 2> gpurun_out/r2fin_detect_config4.err; echo "config4 rc=$?"
python -c "
import json
for l in open('gpurun_out/r2fin_detect_config4.jsonl'):
    d=json.loads(l); print(d['workload'], 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'path', d['path'], 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'], 'cpu', round(d['cpu_baseline']['value']/1e6,1))"
timeout 900 python tools/bench_collect.py > gpurun_out/r2fin_collect_config3.jsonl 2> gpurun_out/r2fin_collect_config3.err; echo "config3 rc=$?"
python -c "
import json
for l in open('gpurun_out/r2fin_collect_config3.jsonl'):
    d=json.loads(l); print(d.get('workload'), 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'], 'frac', d.get('roofline',{}).get('frac'))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect --launch-skip 4 -c 2 -f -o gpurun_out/r2fin_collect python tools/bench_collect.py --workloads shirts_full --queries 2000000 --reps 1 --no-check > gpurun_out/r2fin_ncu_collect.log 2>&1; echo "ncu collect rc=$?"
