set -x
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_encoder.py tests/test_gpu_golden.py tests/test_gpu_embedding.py -x -q 2>&1 | tail -3
timeout 300 python tools/full_phases.py large-v3 120 3 2>&1 | grep -E "step [12]" | sed "s/^/bn256 /"
WDR_GEMM_NO_BN256=1 timeout 300 python tools/full_phases.py large-v3 120 3 2>&1 | grep -E "step [12]" | sed "s/^/bn128 /"
python bench.py --workload encoder --steps 30 --no-cpu-baseline 2>/dev/null | cut -c1-300
WDR_GEMM_NO_BN256=1 python bench.py --workload encoder --steps 30 --no-cpu-baseline 2>/dev/null | cut -c1-300
