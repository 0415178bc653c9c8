#!/bin/bash
# Round 2, GPU round trip L (8 GPUs of one box): update latency over NCCL, bench N=8 weak / strong, config 3 and config 5
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2l_topo.txt 2>&1
timeout 300 python tools/wc_probe.py 8 > gpurun_out/r2l_wc_probe.txt 2>&1; tail -4 gpurun_out/r2l_wc_probe.txt
timeout 300 $TR --master-port 29521 tools/update_probe_nccl.py swim_half > gpurun_out/r2l_update_probe_n8.json 2> gpurun_out/r2l_update_probe.err; echo "update probe rc=$?"; grep "^{" gpurun_out/r2l_update_probe_n8.json | cut -c1-900; tail -2 gpurun_out/r2l_update_probe.err
timeout 500 $TR --master-port 29522 bench.py --gpus 8 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r2l_bench_n8.err
timeout 400 $TR --master-port 29523 bench.py --gpus 8 --scaling strong --no-strong > gpurun_out/r2l_bench_n8_strong.json 2> gpurun_out/r2l_bench_n8_strong.err; echo "bench strong rc=$?"
timeout 600 $TR --master-port 29524 tools/bench_collect.py --gpus 8 > gpurun_out/r2l_collect_config3_n8.jsonl 2> gpurun_out/r2l_collect_n8.err; echo "collect n8 rc=$?"; tail -2 gpurun_out/r2l_collect_n8.err
timeout 600 python tools/bench_collect.py --pool 8 > gpurun_out/r2l_collect_config3_pool8.jsonl 2> gpurun_out/r2l_collect_pool8.err; echo "collect pool8 rc=$?"; tail -2 gpurun_out/r2l_collect_pool8.err
timeout 900 $TR --master-port 29525 tools/bench_mspp.py --gpus 8 --lookahead 1 > gpurun_out/r2l_config5_mspp_n8.jsonl 2> gpurun_out/r2l_mspp_n8.err; echo "mspp n8 rc=$?"; tail -2 gpurun_out/r2l_mspp_n8.err
timeout 600 python tools/bench_mspp.py --lookahead 1 --queries-per-gpu 12582912 > gpurun_out/r2l_config5_mspp_n1.jsonl 2> gpurun_out/r2l_mspp_n1.err; echo "mspp n1 rc=$?"
python - <<'PY'
import json
def load(f):
    try: return [json.loads(l) for l in open(f) if l.startswith('{')]
    except Exception as ex: print(f, ex); return []
for d in load('gpurun_out/r2l_bench_n8.json'):
    e=d['e2e']; print('N8 weak value',d['value']/1e9,'e2e idx',e['value']/1e9,'composed',e['composed_input']['value']/1e9,'lattice',e['lattice_input']['value']/1e9)
    s=d['strong']; print(' pool strong idx', s['value']/1e9, s['ms_per_batch'], 'serial', s['one_batch_at_a_time']['ms_per_batch'], 'composed', s['composed_input']['value']/1e9, s['same_verdicts_as_single_engine'])
for d in load('gpurun_out/r2l_bench_n8_strong.json'):
    e=d['e2e']; print('N8 torchrun strong value',d['value']/1e9,'ms',d['ms_per_step'],'e2e idx',e['value']/1e9,'serial ms',e['one_batch_at_a_time']['ms_per_step'])
for f in ('gpurun_out/r2l_collect_config3_n8.jsonl','gpurun_out/r2l_collect_config3_pool8.jsonl'):
    for d in load(f): print(f[-24:], d['workload'], 'G q/s', (d['value'] or 0)/1e9, 'e2e', d['e2e']['value']/1e9, 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'])
for f in ('gpurun_out/r2l_config5_mspp_n8.jsonl','gpurun_out/r2l_config5_mspp_n1.jsonl'):
    for d in load(f): print(f[-22:], d['value']/1e9, 'G q/s; queries', d['queries_total'], 'update ms', d['update_ms_each'], d['update_ms_each_slowest_rank'], 'mism', d.get('mismatches_vs_oracle'), d.get('checked_queries'))
PY
