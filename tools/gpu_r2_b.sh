#!/bin/bash
# round 2, GPU call B: full parity suite, tile kernel with cached constants, fused finish / per-SNP kernel at
# multi-GPU shard sizes (one GPU emulating one rank's shard), convergence-leg kernel profile
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_pytest.log 2>&1
tail -8 gpurun_out/r2b_pytest.log
for c in 0 1; do timeout 900 python tools/snp_bench.py --cases 3x87,5x256,2x582,1x40 --reps 5 --cache $c 2>&1 | grep -v Warning; done > gpurun_out/r2b_snp.log 2>&1
cat gpurun_out/r2b_snp.log
# shard emulation: 1/8 and 1/4 of C2 on one GPU
for cfg in "150000 212" "300000 425" "1200000 1700"; do set -- $cfg
  for opt in "ld_fused_finish=0" "ld_fused_finish=1"; do
    echo "== snps=$1 blocks=$2 $opt"
    VILMA_B200_OPTIONS=$opt timeout 600 python bench.py --snps $1 --blocks $2 --steps 20 --warmup 5 --extra-workloads none --no-cpu --converge 0 2>&1 >/tmp/o.json | grep "native loop"
    python -c "import json;d=json.load(open('/tmp/o.json'));r=d['roofline'];print('value %.4g ms/step %.4f ld %.4f snp %.4f fin %.4f book %.4f' % (d['value'],d['ms_per_step'],r['ld_kernel']['avg_launch_ms'],r['snp_kernel_avg_ms'],r['finish_kernel_avg_ms'],r['bookkeeping_ms_per_step']))"
  done
done > gpurun_out/r2b_shard.log 2>&1
for opt in "snp_tile=2" "snp_tile=4" "snp_tile=8"; do
  echo "== snps=150000 $opt"
  VILMA_B200_OPTIONS=$opt timeout 600 python bench.py --snps 150000 --blocks 212 --steps 20 --warmup 5 --extra-workloads none --no-cpu --converge 0 2>/dev/null >/tmp/o.json
  python -c "import json;d=json.load(open('/tmp/o.json'));r=d['roofline'];print('value %.4g ms/step %.4f ld %.4f snp %.4f fin %.4f' % (d['value'],d['ms_per_step'],r['ld_kernel']['avg_launch_ms'],r['snp_kernel_avg_ms'],r['finish_kernel_avg_ms']))"
done >> gpurun_out/r2b_shard.log 2>&1
cat gpurun_out/r2b_shard.log
timeout 900 python bench.py --steps 20 --warmup 5 --extra-workloads none --no-cpu --profile-convergence > gpurun_out/r2b_bench_c2_prof.json 2> gpurun_out/r2b_bench_c2_prof.err
python -c "import json;d=json.load(open('gpurun_out/r2b_bench_c2_prof.json'));print(d['value'],d['ms_per_step'],d['convergence'])"
