#!/bin/bash
# Round 2, GPU round trip AC (1 GPU): stage-2 burst chosen by the layout's mean edge run (> 6 edges per entry): 8 (libjagua_cuda.so) / 4 (alt) / 16 (alt2) tests per round
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libjagua_cuda.so libjagua_cuda_alt.so libjagua_cuda_alt2.so; do
  echo "== $lib swim"; JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 1048576 4194304 2>&1 | tee -a gpurun_out/r2ac_size_sweep_swim_$lib.txt
  for w in gardeyn4c_half gardeyn1c_half baldacci3_bin; do
  echo "== $lib $w"; JG_SWEEP_WORKLOAD=$w JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 1048576 2>&1 | tee -a gpurun_out/r2ac_size_sweep_${w}_$lib.txt
  done
done; done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ac_pytest_gpu.log
