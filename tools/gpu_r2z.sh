#!/bin/bash
# Round 2, GPU round trip Z (1 GPU): stage 1 claims and loads its next batch one batch ahead (libjagua_cuda.so; alt2: the same at 4 CTAs/SM, 64 registers) vs alt (no prefetch)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log
for rep in 1 2; do
for lib in libjagua_cuda.so libjagua_cuda_alt.so libjagua_cuda_alt2.so; do
  echo "== $lib"; JG_LIB_NAME=$lib timeout 300 python tools/size_sweep.py 65536 262144 1048576 4194304 2>&1 | tee -a gpurun_out/r2z_size_sweep_$lib.txt
done; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s1_grid --launch-skip 6 -c 1 -f -o gpurun_out/r2z_s1 python tools/size_sweep.py 1048576 > gpurun_out/r2z_ncu_s1.log 2>&1; echo "ncu s1 rc=$?"
