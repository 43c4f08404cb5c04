#!/bin/bash
# final sanity of the driver's default command after the last bench.py edit
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -3 gpurun_out/r2n_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print('C2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_trial_frac'], d['e2e']['value'])
print('conv', d.get('convergence',{}).get('seconds'), 'conv20k', d.get('convergence_20k',{}).get('seconds'))
for w,v in (d.get('workloads') or {}).items():
    print(w, {k:v.get(k) for k in ('value','ms_per_step','skipped','failed')}, v.get('roofline',{}).get('whole_trial_frac'))
    print('   conv', (v.get('convergence') or {}).get('seconds'), (v.get('convergence') or {}).get('iterations')); print('   ckpt', v.get('checkpoint_resume'))
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
