set -x
(timeout 500 python -m pytest tests/test_engine_parity.py tests/test_bench_shape_parity.py -q -m gpu -x -k "latent or flights" > gpurun_out/t_r2p.log 2>&1; tail -3 gpurun_out/t_r2p.log | cut -c1-800)
(timeout 400 python tests/tools/time_latent_h1m.py 2>&1 | tail -14 | grep -v counters | cut -c1-200)
(timeout 600 python scripts/run_h1m_init.py --rows 100000 > gpurun_out/h100k_init_r2p.json 2> gpurun_out/h100k_init_r2p.err; tail -c 1500 gpurun_out/h100k_init_r2p.json; tail -2 gpurun_out/h100k_init_r2p.err | cut -c1-300)
(timeout 900 python bench.py --workload r10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2p_r10m.json 2> gpurun_out/bench_r2p_r10m.err; tail -3 gpurun_out/bench_r2p_r10m.err | cut -c1-400; cut -c1-700 gpurun_out/bench_r2p_r10m.json)
