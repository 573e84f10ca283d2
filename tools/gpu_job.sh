set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python bench.py --workload encoder > gpurun_out/bench_enc.json 2> gpurun_out/bench_enc.err
python bench.py --workload diarize > gpurun_out/bench_diar.json 2> gpurun_out/bench_diar.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_full_ref.json 2> gpurun_out/bench_full_ref.err
cat gpurun_out/bench_full.json | cut -c1-600
