set -x
(timeout 900 python -m pytest tests/test_engine_parity.py -q -m gpu -x > gpurun_out/t_r2m.log 2>&1; tail -5 gpurun_out/t_r2m.log | cut -c1-1500)
(timeout 400 python tests/tools/ab_kblock.py 1000000 gpurun_out/ab_r2m.jsonl 2>&1 | cut -c1-260 | head -8)
(timeout 900 python -m pytest tests/test_bench_shape_parity.py -q -m gpu -x > gpurun_out/t_r2m2.log 2>&1; tail -5 gpurun_out/t_r2m2.log | cut -c1-1500)
