set -x
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q 2>&1 | tail -3
for occ in 3 4 5; do
  WDR_CROSS_OCC=$occ timeout 300 python tools/full_phases.py large-v3 120 3 2>&1 | grep -E "step [12]" | sed "s/^/online occ$occ /" | tee -a gpurun_out/phases_occ.log
done
WDR_CROSS_TWO_PASS=1 timeout 300 python tools/full_phases.py large-v3 120 3 2>&1 | grep -E "step [12]" | sed "s/^/twopass occ6 /" | tee -a gpurun_out/phases_occ.log
