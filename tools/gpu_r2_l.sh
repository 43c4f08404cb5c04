#!/bin/bash
# round 2, GPU call L (one GPU): final build -- parity suite, the driver's two bench commands, launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2l_pytest.log 2>&1
tail -4 gpurun_out/r2l_pytest.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -3 gpurun_out/r2l_bench.err | cut -c1-300
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.err
cat gpurun_out/r2l_bench_reference.json | cut -c1-600
B="python bench.py --steps 3 --warmup 3 --extra-workloads none --no-cpu --converge 0 --no-checkpoint-leg"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vb_(ld_sym|ld_finish|ld_matvec|ld_fac|snp|sum_|pm_diff|stats|scale|init)" -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps3_warmup3.csv $B > /dev/null 2> gpurun_out/r2l_ncu_launch.err
wc -l gpurun_out/r02_launches_bench_steps3_warmup3.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vb_ld_fac_kernel -s 6 -c 1 -o /tmp/r02_ld_fac_full -f python bench.py --workload c3 --steps 3 --warmup 3 --extra-workloads none --no-cpu --converge 0 --no-checkpoint-leg > gpurun_out/r02_ld_fac_full.log 2>&1
ncu -i /tmp/r02_ld_fac_full.ncu-rep --page raw --csv > gpurun_out/r02_ld_fac_full_raw.csv 2>/dev/null
ncu -i /tmp/r02_ld_fac_full.ncu-rep --page details > gpurun_out/r02_ld_fac_full_details.txt 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print('C2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_trial_frac'], d['e2e']['value'])
print('conv', d.get('convergence'))
print('cpu', d.get('cpu_baseline'))
print('conv20k', d.get('convergence_20k'))
for w,v in (d.get('workloads') or {}).items():
    print(w, {k:v.get(k) for k in ('value','ms_per_step','skipped','failed')}, v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('whole_trial_frac'), v.get('roofline',{}).get('kernel'))
    print('   conv', v.get('convergence')); print('   ckpt', v.get('checkpoint_resume')); print('   cpu', v.get('cpu_baseline'))
PY
