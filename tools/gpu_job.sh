set -x
timeout 900 python -m pytest tests/test_gpu_embedding.py -x -q > gpurun_out/t_emb.log 2>&1; echo "rc=$?" >> gpurun_out/t_emb.log
tail -15 gpurun_out/t_emb.log
WDR_LANES=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ph1.json 2> gpurun_out/bench_ph1.err
WDR_LANES=2 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ph2.json 2> gpurun_out/bench_ph2.err
python - <<'PY'
import json
for f in ("ph1","ph2"):
    d=json.load(open(f"gpurun_out/bench_{f}.json"))
    print(f, round(d["ms_per_step"]), round(d["e2e"]["ms_per_step"]), d["phases_ms_last_step"])
PY
