#!/bin/bash
# Round 2, GPU round trip K (1 GPU): smaller stage 2 (rolled quadrant test), staged split variants; A/B of JG_STAGE_S2
mkdir -p gpurun_out
JG_STAGE_S2=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_staged.log 2>&1; echo "pytest(staged, all) rc=$?"; tail -3 gpurun_out/r2k_pytest_staged.log
for v in 0 1; do
  echo "JG_STAGE_S2=$v"; JG_STAGE_S2=$v timeout 200 python tools/size_sweep.py 16384 65536 262144 1048576 1048576 4194304 > gpurun_out/r2k_size_sweep_stage$v.txt 2>&1; cat gpurun_out/r2k_size_sweep_stage$v.txt
done
for v in 0 1; do
JG_STAGE_S2=$v timeout 300 python bench.py --no-cpu-baseline --no-strong > gpurun_out/r2k_bench_stage$v.json 2> gpurun_out/r2k_bench_stage$v.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2k_bench_stage$v.json') if l.startswith('{')][0]; e=d['e2e']
print('stage $v: value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms idx', round(e['one_batch_at_a_time']['ms_per_step'],3), 'lattice', round(e['lattice_input']['value']/1e9,2))"
done
