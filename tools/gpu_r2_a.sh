#!/bin/bash
# round 2, GPU call A: parity suite on the new build, tile-kernel variants, a quick C2 bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g | head -2 >> gpurun_out/r2a_gpu.txt; df -h /dev/shm /tmp >> gpurun_out/r2a_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
for v in base old u1 t384 t512 pair pair384 noprec noregpf; do
  echo "== $v"
  VILMA_B200_LIB=variants/lib_$v.so timeout 600 python tools/snp_bench.py --cases 3x87,5x256,2x582 --reps 5 2>&1 | grep -v Warning
done > gpurun_out/r2a_snp.log 2>&1
cat gpurun_out/r2a_snp.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --extra-workloads none --no-cpu ) > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
tail -3 gpurun_out/r2a_bench_c2.err; head -c 1500 gpurun_out/r2a_bench_c2.json
