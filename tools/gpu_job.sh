set -x
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q 2>&1 | tail -3
WDR_DEBUG_TIMING=1 timeout 300 python tools/full_phases.py large-v3 120 4 2>&1 | grep -E "step|wdr" | tee gpurun_out/phases.log
