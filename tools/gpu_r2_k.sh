#!/bin/bash
# round 2, GPU call K (one GPU): parity suite with the read-once factor kernel, C3 with / without it,
# multi-GPU flavour of the per-SNP kernel (fused annotation sums) at a 1/8 shard: slots vs shuffles + parking
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2k_pytest.log 2>&1
tail -6 gpurun_out/r2k_pytest.log
for o in 1; do
  VILMA_B200_OPTIONS=ld_factor_once=$o timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --extra-workloads none --no-cpu --converge 0 --no-checkpoint-leg > gpurun_out/r2k_c3_once$o.json 2> gpurun_out/r2k_c3_once$o.err
  python -c "
import json
d=json.load(open('gpurun_out/r2k_c3_once$o.json')); r=d['roofline']
print('c3 once=$o', d['value'], d['ms_per_step'], r['frac'], r.get('ld_kernel'), r.get('snp_kernel_avg_ms'), r.get('whole_trial_frac'), d['config'].get('ld_store'))"
done
