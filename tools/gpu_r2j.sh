#!/bin/bash
# Round 2, GPU round trip J (1 GPU): stage 2 with the layout's entries + edges staged by TMA bulk copies (JG_STAGE_S2=1) vs L1
mkdir -p gpurun_out
JG_STAGE_S2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "detect_matches_oracle or poly_only or multi_segment or filters or boundaries or near_item or self_key" > gpurun_out/r2j_pytest_staged.log 2>&1; echo "pytest(staged) rc=$?"; tail -3 gpurun_out/r2j_pytest_staged.log
for v in 0 1; do
  echo "JG_STAGE_S2=$v"; JG_STAGE_S2=$v timeout 200 python tools/size_sweep.py 65536 262144 1048576 1048576 > gpurun_out/r2j_size_sweep_stage$v.txt 2>&1; cat gpurun_out/r2j_size_sweep_stage$v.txt
done
JG_STAGE_S2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s2_grid --launch-skip 12 -c 1 -f -o gpurun_out/r2j_s2_staged python tools/size_sweep.py 1048576 > gpurun_out/r2j_ncu_staged.log 2>&1; echo "ncu staged rc=$?"
JG_STAGE_S2=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s2_grid --launch-skip 12 -c 1 -f -o gpurun_out/r2j_s2_l1 python tools/size_sweep.py 1048576 > gpurun_out/r2j_ncu_l1.log 2>&1; echo "ncu l1 rc=$?"
