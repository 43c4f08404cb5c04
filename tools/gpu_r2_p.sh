#!/bin/bash
# multi-GPU: sharded upload A/B -- zero-copy gather by the GPU against host staging (public call, e2e breakdown)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -k "not snp_tile" ) > gpurun_out/r2p_multi_n$N.log 2>&1
tail -4 gpurun_out/r2p_multi_n$N.log
for z in 1 0; do
( VILMA_B200_OPTIONS=shard_zero_copy=$z timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$z \
    bench.py --gpus $N --steps 20 --warmup 5 --extra-workloads none --no-cpu --converge 0 ) > gpurun_out/r2p_bench_n${N}_z$z.json 2> gpurun_out/r2p_bench_n${N}_z$z.err
echo "zero_copy=$z"; grep -o "\[rank [0-9]\] e2e[^\[]*" gpurun_out/r2p_bench_n${N}_z$z.err
done
