set -x
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
timeout 300 python tools/full_phases.py large-v3 120 3 2>&1 | grep -E "step [12]" | sed "s/^/kbmajor /"
