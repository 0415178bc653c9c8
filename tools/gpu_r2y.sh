#!/bin/bash
# Round 2, GPU round trip Y (1 GPU): stage 1 phase A in two steps (libjagua_cuda.so) vs one pass over all poles (alt, -DG1_SPLIT_A=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_pytest_gpu.log
for rep in 1 2; do
for lib in libjagua_cuda.so libjagua_cuda_alt.so; do
  echo "== $lib"; JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 65536 262144 1048576 4194304 2>&1 | tee -a gpurun_out/r2y_size_sweep_$lib.txt
done; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s1_grid --launch-skip 6 -c 1 -f -o gpurun_out/r2y_s1 python tools/size_sweep.py 1048576 > gpurun_out/r2y_ncu_s1.log 2>&1; echo "ncu s1 rc=$?"
