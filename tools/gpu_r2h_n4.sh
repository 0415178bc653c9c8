#!/bin/bash
# Round 2, GPU round trip H (4 GPUs of one box): pool + NCCL-broadcast tests, WC probe, bench N=4 weak / strong, config 3 and 5
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2h_pytest_n4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest_n4.log
timeout 120 python tools/wc_probe.py 4 > gpurun_out/r2h_wc_probe.txt 2>&1; cat gpurun_out/r2h_wc_probe.txt
timeout 400 $TR --master-port 29511 bench.py --gpus 4 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err; echo "bench n4 rc=$?"; tail -3 gpurun_out/r2h_bench_n4.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2h_bench_n4.json'))
    e=d['e2e']; print('N4 value',d['value']/1e9,'e2e idx',e['value']/1e9,'composed',e['composed_input']['value']/1e9,'lattice',e['lattice_input']['value']/1e9)
    s=d['strong']; print('strong idx', s['value']/1e9, s['ms_per_batch'], 'serial', s['one_batch_at_a_time']['ms_per_batch'], 'composed', s['composed_input']['value']/1e9, s['same_verdicts_as_single_engine'])
    print(d['layout_broadcast'])
except Exception as ex: print('parse failed', ex)
PY
timeout 400 $TR --master-port 29512 bench.py --gpus 4 --scaling strong --no-strong > gpurun_out/r2h_bench_n4_strong.json 2> gpurun_out/r2h_bench_n4_strong.err; echo "bench strong rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_n4_strong.json')); e=d['e2e']; print('STRONG torchrun value',d['value']/1e9,'ms',d['ms_per_step'],'e2e idx',e['value']/1e9,'serial ms',e['one_batch_at_a_time']['ms_per_step'])"
timeout 600 python tools/bench_collect.py --pool 4 > gpurun_out/r2h_collect_config3_pool4.jsonl 2> gpurun_out/r2h_collect_pool4.err; echo "collect pool rc=$?"; tail -2 gpurun_out/r2h_collect_pool4.err
timeout 600 $TR --master-port 29513 tools/bench_collect.py --gpus 4 > gpurun_out/r2h_collect_config3_n4.jsonl 2> gpurun_out/r2h_collect_n4.err; echo "collect n4 rc=$?"; tail -2 gpurun_out/r2h_collect_n4.err
python - <<'PY'
import json
for f in ('gpurun_out/r2h_collect_config3_pool4.jsonl','gpurun_out/r2h_collect_config3_n4.jsonl'):
    try:
        for ln in open(f):
            d=json.loads(ln); print(f[-22:], d['workload'], 'G q/s', (d['value'] or 0)/1e9, 'e2e', d['e2e']['value']/1e9, 'checked', d['checked_queries'], 'mism', d['mismatches_vs_oracle'])
    except Exception as ex: print(f, 'failed', ex)
PY
timeout 600 $TR --master-port 29514 tools/bench_mspp.py --gpus 4 --queries-per-gpu 6291456 --lookahead 1 > gpurun_out/r2h_config5_mspp_n4.jsonl 2> gpurun_out/r2h_mspp_n4.err; echo "mspp rc=$?"; tail -2 gpurun_out/r2h_mspp_n4.err
python -c "
import json; d=json.loads(open('gpurun_out/r2h_config5_mspp_n4.jsonl').read()); print('MSPP N4', d['value']/1e9, 'update ms', d['update_ms_each'], d['update_ms_each_slowest_rank'], 'mism', d.get('mismatches_vs_oracle'), d.get('checked_queries'))"
