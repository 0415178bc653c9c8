#!/bin/bash
# Round 2, GPU round trip R (1 GPU): stage 2 with convergent rounds (libjagua_cuda.so) vs the nested loops (libjagua_cuda_alt.so, -DJG_S2_CONV=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest_gpu.log
for rep in 1 2; do
for lib in libjagua_cuda.so libjagua_cuda_alt.so; do
  echo "== $lib"; JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 65536 262144 1048576 4194304 2>&1 | tee -a gpurun_out/r2r_size_sweep_$lib.txt
done; done
for lib in libjagua_cuda.so libjagua_cuda_alt.so; do
  JG_LIB_NAME=$lib timeout 300 python bench.py --no-cpu-baseline --no-strong > gpurun_out/r2r_bench_$lib.json 2> gpurun_out/r2r_bench_$lib.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2r_bench_$lib.json').read()); e=d['e2e']
print('$lib value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],4), 'e2e idx', round(e['value']/1e9,2), 'serial ms', round(e['one_batch_at_a_time']['ms_per_step'],3))"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s2_grid --launch-skip 6 -c 1 -f -o gpurun_out/r2r_s2_conv python tools/size_sweep.py 1048576 > gpurun_out/r2r_ncu_s2.log 2>&1; echo "ncu s2 rc=$?"
