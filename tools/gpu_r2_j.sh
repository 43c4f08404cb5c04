#!/bin/bash
# round 2, GPU call J (one GPU): ncu evidence on the final build -- launch list of the bench command, --set full
# of the dominant kernels (exported to CSV / text on the box; the .ncu-rep files are too large to bring back)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --extra-workloads none --no-cpu --converge 0 --no-checkpoint-leg"
$B > gpurun_out/r2j_bench_plain.json 2> gpurun_out/r2j_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/r2j_bench_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vb_(ld_sym|ld_finish|ld_matvec|ld_fac|snp|sum_|pm_diff|stats|scale|init)" -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps3_warmup3.csv $B > /dev/null 2> gpurun_out/r2j_ncu_launch.err
wc -l gpurun_out/r02_launches_bench_steps3_warmup3.csv
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o /tmp/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ls -la /tmp/$name.ncu-rep | awk '{print $5, $9}'
}
cap r02_ld_sym_full vb_ld_sym_kernel 6 1 $B
cap r02_snp3_full vb_snp3_kernel 6 2 $B
cap r02_finish_full vb_ld_finish_sym_kernel 6 1 $B
cap r02_ld_sym_slabs_full vb_ld_sym_kernel 6 1 python bench.py --snps 600000 --blocks 170 --steps 3 --warmup 3 --extra-workloads none --no-cpu --converge 0 --no-checkpoint-leg
cap r02_tile_p3_k87_full vb_snp_tile 6 2 python tools/snp_bench.py --cases 3x87 --reps 2
cap r02_tile_p5_k256_full vb_snp_tile 6 2 python tools/snp_bench.py --cases 5x256 --reps 2
python bench.py --snps 600000 --blocks 170 --steps 10 --warmup 3 --extra-workloads none --no-cpu --converge 0 --no-checkpoint-leg > gpurun_out/r2j_bench_slabs.json 2>/dev/null
python -c "
import json
for f in ('r2j_bench_plain','r2j_bench_slabs'):
    d=json.load(open('gpurun_out/%s.json'%f)); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['ld_kernel'], d['config'].get('ld_store'))
"
du -sh gpurun_out
