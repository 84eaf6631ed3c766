set -x
(timeout 1500 python -m pytest tests/ -q -m gpu -s > gpurun_out/t_r2o.log 2>&1; grep -a "^F1\|passed\|failed\|slots \|appended" gpurun_out/t_r2o.log | cut -c1-400; tail -5 gpurun_out/t_r2o.log | cut -c1-800)
(timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; tail -2 gpurun_out/bench_r2o.err | cut -c1-400; cut -c1-900 gpurun_out/bench_r2o.json)
(timeout 400 python bench.py --sweep all --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2o_all.json 2> gpurun_out/bench_r2o_all.err; tail -2 gpurun_out/bench_r2o_all.err | cut -c1-400; cut -c1-500 gpurun_out/bench_r2o_all.json)
(timeout 700 python scripts/run_h1m_init.py --rows 1000000 > gpurun_out/h1m_init_r2o.json 2> gpurun_out/h1m_init_r2o.err; tail -c 1800 gpurun_out/h1m_init_r2o.json; tail -2 gpurun_out/h1m_init_r2o.err | cut -c1-300)
