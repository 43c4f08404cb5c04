set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t4.log 2>&1; tail -5 gpurun_out/t4.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c2_d.json 2> gpurun_out/bench_c2_d.err; tail -3 gpurun_out/bench_c2_d.err; cat gpurun_out/bench_c2_d.json
python tools/snp_bench.py --cases 1x14,2x42,2x582,3x123,5x256 > gpurun_out/sb4_auto.log 2>&1; cat gpurun_out/sb4_auto.log
python tools/snp_bench.py --cases 1x14 --snps 150000 > gpurun_out/sb4_small.log 2>&1; cat gpurun_out/sb4_small.log
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_c.json 2> gpurun_out/bench_c3_c.err; tail -4 gpurun_out/bench_c3_c.err; cat gpurun_out/bench_c3_c.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vb_ld_finish -s 6 -c 2 -o gpurun_out/prof_finish -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_fin.log 2>&1; tail -3 gpurun_out/ncu_fin.log
