set -x
(timeout 600 python -m pytest tests/test_engine_parity.py tests/test_bench_shape_parity.py -q -m gpu -x -k "latent or compaction or flights" > gpurun_out/t_r2n.log 2>&1; tail -5 gpurun_out/t_r2n.log | cut -c1-1500)
(timeout 400 python tests/tools/time_latent_h1m.py 2>&1 | tail -16 | cut -c1-200)
