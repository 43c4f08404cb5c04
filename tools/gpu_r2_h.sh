#!/bin/bash
# multi-GPU validation (run with gpurun --gpus N): sharded parity tests, then the driver's bench command at N
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -q ) > gpurun_out/r2h2_multi_n$N.log 2>&1
tail -5 gpurun_out/r2h2_multi_n$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 20 --warmup 5 $EXTRA ) > gpurun_out/r2h2_bench_n$N.json 2> gpurun_out/r2h2_bench_n$N.err
grep -E "native loop|e2e|built|failed|Error" gpurun_out/r2h2_bench_n$N.err | tail -30
python - <<PY
import json
d=json.load(open('gpurun_out/r2h2_bench_n$N.json'))
print('C2 N=$N', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds'])
print('conv', {k:v for k,v in d.get('convergence',{}).items() if k in ('seconds','iterations','trials','final_elbo')})
print('per rank', d['roofline']['per_rank_ld_ms'], d['roofline']['per_rank_snp_ms'], d['roofline']['finish_kernel_avg_ms'])
for w,v in (d.get('workloads') or {}).items():
    print(w, {k:v.get(k) for k in ('value','ms_per_step','skipped','failed')}, v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('whole_trial_frac'))
    print('   conv', v.get('convergence')); print('  cfg', v.get('config',{}).get('ld_store'), v.get('config',{}).get('setup_s'))
PY
