#!/bin/bash
# Round 2, GPU round trip D (1 GPU): the lane = (query, terminal) collect kernel - parity, timing, config 3, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "collect or pool or seeded or piers or layout_updates or golden or smoke" > gpurun_out/r2d_pytest_collect.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest_collect.log
timeout 120 python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2d_collect_shirts.txt 2>&1; tail -9 gpurun_out/r2d_collect_shirts.txt
timeout 120 python tools/collect_probe.py trousers_full 1048576 > gpurun_out/r2d_collect_trousers.txt 2>&1; tail -1 gpurun_out/r2d_collect_trousers.txt
timeout 120 python tools/collect_probe.py swim_half 262144 > gpurun_out/r2d_collect_swim.txt 2>&1; tail -1 gpurun_out/r2d_collect_swim.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect_grid --launch-skip 3 -c 1 -f -o gpurun_out/r2d_collect_shirts python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2d_ncu_collect.log 2>&1; echo "ncu rc=$?"
