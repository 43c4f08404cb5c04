set -x
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/t8_multi.log 2>&1; tail -3 gpurun_out/t8_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_b.json 2> gpurun_out/bench_n2_b.err; tail -4 gpurun_out/bench_n2_b.err; cat gpurun_out/bench_n2_b.json
