#!/bin/bash
# Round 2, GPU round trip F (1 GPU): collect kernels with integer candidate windows / early skips
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "collect or pool or seeded or piers or layout_updates or golden or smoke" > gpurun_out/r2f_pytest_collect.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest_collect.log
timeout 120 python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2f_collect_shirts.txt 2>&1; tail -9 gpurun_out/r2f_collect_shirts.txt
timeout 120 python tools/collect_probe.py trousers_full 1048576 > gpurun_out/r2f_collect_trousers.txt 2>&1; tail -1 gpurun_out/r2f_collect_trousers.txt
timeout 120 python tools/collect_probe.py swim_half 262144 > gpurun_out/r2f_collect_swim.txt 2>&1; tail -1 gpurun_out/r2f_collect_swim.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect --launch-skip 6 -c 2 -f -o gpurun_out/r2f_collect_shirts python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2f_ncu_collect.log 2>&1; echo "ncu rc=$?"
