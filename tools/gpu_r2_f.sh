#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for m in 75000 300000 1200000 2400000; do
  timeout 600 python tools/snp_bench.py --cases 3x87,5x256 --reps 5 --cache 0 --snps $m --opt snp_tile_ring=0 2>&1 | grep -v Warning
done > gpurun_out/r2f_snp.log 2>&1
cat gpurun_out/r2f_snp.log
