#!/bin/bash
# Round 2, GPU round trip U (1 GPU): per-piece device timeline of one 1M indexed batch under the library's schedule and under fixed pieces
mkdir -p gpurun_out
for cfg in "0 4" "262144 4" "1048576 3" "524288 4"; do
  echo "== chunk/streams: $cfg (indexed)"; JG_TRACE=1 timeout 120 python tools/trace_e2e.py $cfg indexed 2>&1 | tail -14
done 2>&1 | tee gpurun_out/r2u_trace.txt
timeout 600 python tools/schedule_probe.py 1048576 2097152 2>&1 | tee gpurun_out/r2u_schedule_probe.txt
