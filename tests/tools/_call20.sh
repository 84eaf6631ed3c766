set -x
(PCLEAN_LOG_ALLOC=1 timeout 600 python bench.py --workload r10m --rows 1000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2r_r1m.json 2> gpurun_out/bench_r2r_r1m.err; grep -a "allocation\|bench r0\|Error" gpurun_out/bench_r2r_r1m.err | cut -c1-300 | tail -40; cut -c1-700 gpurun_out/bench_r2r_r1m.json)
