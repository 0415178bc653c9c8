#!/bin/bash
# Round 2, GPU round trip A (1 GPU): parity tests, bench line (+ reference arm), collect timing + ncu source profile, WC probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
timeout 120 python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2a_collect_shirts.txt 2>&1; tail -2 gpurun_out/r2a_collect_shirts.txt
timeout 120 python tools/collect_probe.py trousers_full 1048576 > gpurun_out/r2a_collect_trousers.txt 2>&1; tail -1 gpurun_out/r2a_collect_trousers.txt
timeout 60 python tools/wc_probe.py 1 > gpurun_out/r2a_wc_probe.txt 2>&1; cat gpurun_out/r2a_wc_probe.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_collect_grid --launch-skip 3 -c 1 -f -o gpurun_out/r2a_collect_shirts python tools/collect_probe.py shirts_full 1048576 > gpurun_out/r2a_ncu_collect.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | head -30
