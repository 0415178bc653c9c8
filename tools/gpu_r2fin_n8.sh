#!/bin/bash
# Round 2, final state on 8 GPUs of one box: bench N=8 weak (with the pool's strong leg), bench N=8 strong, config 3 at N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29522 bench.py --gpus 8 > gpurun_out/r2fin_bench_n8.json 2> gpurun_out/r2fin_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r2fin_bench_n8.err
timeout 200 $TR --master-port 29526 bench.py --gpus 8 --impl reference > gpurun_out/r2fin_bench_n8_ref.json 2> gpurun_out/r2fin_bench_n8_ref.err; echo "bench n8 ref rc=$?"
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --scaling strong --no-strong --no-cpu-baseline > gpurun_out/r2fin_bench_n8_strong.json 2> gpurun_out/r2fin_bench_n8_strong.err; echo "bench strong rc=$?"
timeout 500 $TR --master-port 29524 tools/bench_collect.py --gpus 8 > gpurun_out/r2fin_collect_config3_n8.jsonl 2> gpurun_out/r2fin_collect_n8.err; echo "collect n8 rc=$?"; tail -2 gpurun_out/r2fin_collect_n8.err
python - <<'PY'
import json
def load(f):
    try: return [json.loads(l) for l in open(f) if l.startswith('{')]
    except Exception as ex: print(f, ex); return []
for d in load('gpurun_out/r2fin_bench_n8.json'):
    e=d['e2e']; print('N8 weak value',d['value']/1e9,'e2e idx',e['value']/1e9,'composed',e['composed_input']['value']/1e9,'lattice',e['lattice_input']['value']/1e9, 'serial ms', e['one_batch_at_a_time']['ms_per_step'])
    s=d.get('strong') or {}; print(' pool strong', {k:v for k,v in s.items() if k in ('value','ms_per_batch','same_verdicts_as_single_engine','error')})
    print(' bcast', d.get('layout_broadcast'), 'mism', d['cpu_baseline'].get('mismatches_vs_gpu'))
for d in load('gpurun_out/r2fin_bench_n8_ref.json'): print('ref arm', d.get('value'), d.get('cpu_baseline',{}).get('cores'))
for d in load('gpurun_out/r2fin_bench_n8_strong.json'):
    e=d['e2e']; print('N8 torchrun strong value',d['value']/1e9,'ms',d['ms_per_step'],'e2e idx',e['value']/1e9,'serial ms',e['one_batch_at_a_time']['ms_per_step'])
for d in load('gpurun_out/r2fin_collect_config3_n8.jsonl'): print(d['workload'], 'G q/s', (d['value'] or 0)/1e9, 'e2e', d['e2e']['value']/1e9, 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'])
PY
