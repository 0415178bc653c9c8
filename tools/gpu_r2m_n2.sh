#!/bin/bash
# Round 2, GPU round trip M (2 GPUs): pool with the TMA-staged kernels on more than one device; bench N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2m_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest_n2.log
timeout 400 $TR --master-port 29531 bench.py --gpus 2 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/r2m_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2m_bench_n2.json'):
    if not l.startswith('{'): print('EXTRA STDOUT LINE:', l[:80]); continue
    d=json.loads(l); e=d['e2e']; print('N2 weak value',d['value']/1e9,'e2e idx',e['value']/1e9,'composed',e['composed_input']['value']/1e9,'lattice',e['lattice_input']['value']/1e9)
    print(' strong', json.dumps(d['strong'])[:600])
PY
