#!/bin/bash
# multi-GPU: sharded parity + re-upload checks, then the bench (C2 only) for the e2e breakdown
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x ) > gpurun_out/r2i_multi_n$N.log 2>&1
tail -25 gpurun_out/r2i_multi_n$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 20 --warmup 5 --extra-workloads none --no-cpu --converge 0 ) > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err
grep -o "\[rank [0-9]\] e2e[^\[]*" gpurun_out/r2i_bench_n$N.err; tail -3 gpurun_out/r2i_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2i_bench_n$N.json'))
print('C2 N=$N', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds'])
print('conv', {k:v for k,v in d.get('convergence',{}).items() if k in ('seconds','iterations','trials','final_elbo')})
PY
