#!/bin/bash
# Round 2, GPU round trip AB (1 GPU): stage-2 rounds with up to 4 (libjagua_cuda.so) / 1 (alt) / 2 (alt2) edge tests per round, swim and the 1043-vertex gardeyn4_c layout
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libjagua_cuda.so libjagua_cuda_alt.so libjagua_cuda_alt2.so; do
  echo "== $lib swim"; JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 262144 1048576 4194304 2>&1 | tee -a gpurun_out/r2ab_size_sweep_swim_$lib.txt
  echo "== $lib gardeyn4c"; JG_SWEEP_WORKLOAD=gardeyn4c_half JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 262144 1048576 2>&1 | tee -a gpurun_out/r2ab_size_sweep_gardeyn4c_$lib.txt
  echo "== $lib gardeyn1c"; JG_SWEEP_WORKLOAD=gardeyn1c_half JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 1048576 2>&1 | tee -a gpurun_out/r2ab_size_sweep_gardeyn1c_$lib.txt
done; done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ab_pytest_gpu.log
