#!/bin/bash
# multi-GPU: the driver's bench command at N (C2 headline only), twice, for run-to-run spread
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for rep in ${REPS:-a b}; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N \
    bench.py --gpus $N --steps 20 --warmup 5 --extra-workloads none --no-cpu ) > gpurun_out/r2m_bench_n${N}_$rep.json 2> gpurun_out/r2m_bench_n${N}_$rep.err
grep -o "\[rank 0\] e2e[^\[]*" gpurun_out/r2m_bench_n${N}_$rep.err; grep -o "\[rank 0\] native loop[^\[]*" gpurun_out/r2m_bench_n${N}_$rep.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_n${N}_$rep.json'))
print('C2 N=$N', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds'])
print('conv', {k:v for k,v in d.get('convergence',{}).items() if k in ('seconds','iterations','trials','final_elbo')})
print('per rank', d['roofline']['per_rank_ld_ms'], d['roofline']['per_rank_snp_ms'], d['roofline']['finish_kernel_avg_ms'])
PY
done
