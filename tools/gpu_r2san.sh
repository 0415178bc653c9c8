#!/bin/bash
# Round 2, compute-sanitizer on the final build (1 GPU): memcheck over the parity tests (grid + walk paths, staged stage 2 with TMA bulk copies,
# collect with (query, terminal, entry) lanes, patch uploads, pool on one device), racecheck over the kernels that use shared-memory queues
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_ref_case.py tests/test_golden.py -m gpu -x -q -k "not lbf and not nccl and not piece_schedule" > gpurun_out/r2san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2san_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "(test_collect_matches_oracle and (swim_half or shirts_full)) or (test_detect_matches_oracle and swim_half) or (test_collect_multi_matches_oracle and swim_half) or test_queries_follow_layout_updates" > gpurun_out/r2san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2san_racecheck.log | tail -3
