set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload diarize --steps 5 --no-cpu-baseline > gpurun_out/bench_diar_n2.json 2> gpurun_out/bench_diar_n2.err; tail -3 gpurun_out/bench_diar_n2.err
cat gpurun_out/bench_n2.json | cut -c1-400; cat gpurun_out/bench_diar_n2.json | cut -c1-900
