#!/bin/bash
# round 2, GPU call D: tile kernel with the per-warp TMA ring (with / without the cached constants), parity suite
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_pytest.log 2>&1
tail -6 gpurun_out/r2d_pytest.log
for opt in "snp_tile_ring=0" "snp_tile_ring=1"; do for c in 1 0; do
  timeout 900 python tools/snp_bench.py --cases 3x87,5x256,2x582,1x40,4x31 --reps 5 --cache $c --opt $opt 2>&1 | grep -v Warning
done; done > gpurun_out/r2d_snp.log 2>&1
cat gpurun_out/r2d_snp.log
