#!/bin/bash
# round 2, GPU call C: ncu --set full of the tile kernel (P=5 K=256; P=3 K=87) with and without the cache;
# the reports are exported to CSV / text on the box (the .ncu-rep files are too large to bring back)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cap() {  # name, cases, cache
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vb_snp_tile -s 6 -c 2 \
     -o /tmp/$1 -f python tools/snp_bench.py --cases $2 --reps 2 --cache $3 > gpurun_out/$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source_sass.csv 2>/dev/null
  tail -1 gpurun_out/$1.log
}
cap r2c_tile_p5_cache0 5x256 0
cap r2c_tile_p5_cache1 5x256 1
cap r2c_tile_p3_cache1 3x87 1
du -sh gpurun_out
