#!/bin/bash
# Round 2, GPU round trip AL (1 GPU): containment tests of the split path inside k_s2_grid<.., DEFER> (JG_DEFER_CT=1) vs k_s2_contain as a third launch (0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2al_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2al_pytest_gpu.log
for v in 1 0 1 0; do echo "JG_DEFER_CT=$v"; JG_DEFER_CT=$v python tools/size_sweep.py 786432 1048576 4194304; done 2>&1 | tee gpurun_out/r2al_defer_ab.txt
