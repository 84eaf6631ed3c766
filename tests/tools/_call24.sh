set -x
(timeout 900 python -m pytest tests/test_engine_parity.py tests/test_bench_shape_parity.py -q -m gpu -x -k "rents or flights or memo" > gpurun_out/t_r2u.log 2>&1; tail -8 gpurun_out/t_r2u.log | cut -c1-900)
(timeout 600 python bench.py --workload r10m --rows 1000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2u_r1m.json 2> gpurun_out/bench_r2u_r1m.err; grep -a "bench r0\|Error" gpurun_out/bench_r2u_r1m.err | cut -c1-300 | tail -4; cut -c1-330 gpurun_out/bench_r2u_r1m.json)
(timeout 1500 python bench.py --workload r10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2u_r10m.json 2> gpurun_out/bench_r2u_r10m.err; grep -a "bench r0\|Error" gpurun_out/bench_r2u_r10m.err | cut -c1-300 | tail -4; cut -c1-330 gpurun_out/bench_r2u_r10m.json)
(timeout 300 python bench.py --workload rents --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2u_rents.json 2> gpurun_out/bench_r2u_rents.err; cut -c1-330 gpurun_out/bench_r2u_rents.json)
