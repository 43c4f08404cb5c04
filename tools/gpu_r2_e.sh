#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for v in nopf nopf_regpf pf2 pf8; do for c in 0 1; do
  echo "== $v cache=$c"
  VILMA_B200_LIB=variants/lib_$v.so timeout 600 python tools/snp_bench.py --cases 3x87,5x256,2x582 --reps 5 --cache $c --opt snp_tile_ring=0 2>&1 | grep -v Warning
done; done > gpurun_out/r2e_snp.log 2>&1
cat gpurun_out/r2e_snp.log
