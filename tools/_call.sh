set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; tail -5 gpurun_out/t2.log
python tools/snp_bench.py --cases 1x14,2x42,2x582,3x123,5x256 > gpurun_out/sb2_auto.log 2>&1
python tools/snp_bench.py --cases 1x14 --opt snp_tile=0 > gpurun_out/sb2_c2old.log 2>&1
python tools/snp_bench.py --cases 1x14 --opt snp_tile=2 > gpurun_out/sb2_c2w2.log 2>&1
for ua in 1 4; do VILMA_B200_LIB=$PWD/variants/lib_ua$ua.so python tools/snp_bench.py --cases 2x582,3x123 > gpurun_out/sb2_ua$ua.log 2>&1; done
cat gpurun_out/sb2_*.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c2_b.json 2> gpurun_out/bench_c2_b.err; tail -3 gpurun_out/bench_c2_b.err; cat gpurun_out/bench_c2_b.json
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_a.json 2> gpurun_out/bench_c3_a.err; tail -8 gpurun_out/bench_c3_a.err; cat gpurun_out/bench_c3_a.json
