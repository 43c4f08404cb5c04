#!/bin/bash
# GPU call G: parity suite, per-SNP micro-benchmarks, and the driver's default bench command (headline C2 +
# workloads c3/c5 + CPU legs) exactly as the driver runs it
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2g_pytest.log 2>&1
tail -5 gpurun_out/r2g_pytest.log
timeout 900 python tools/snp_bench.py --cases 3x87,5x256,2x582,1x14 --reps 5 2>&1 | grep -v Warning > gpurun_out/r2g_snp.log
cat gpurun_out/r2g_snp.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -25 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('C2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_trial_frac'], d['e2e']['value'])
print('conv', d.get('convergence'))
print('cpu', d.get('cpu_baseline'))
print('conv20k', d.get('convergence_20k'))
for w,v in (d.get('workloads') or {}).items():
    print(w, {k:v.get(k) for k in ('value','ms_per_step','skipped','failed')}, v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('whole_trial_frac'), v.get('roofline',{}).get('kernel'))
    print('   conv', v.get('convergence')); print('   ckpt', v.get('checkpoint_resume')); print('   cpu', v.get('cpu_baseline'))
PY
