set -x
(timeout 900 python bench.py --workload r10m --rows 3000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2q_r3m.json 2> gpurun_out/bench_r2q_r3m.err; tail -3 gpurun_out/bench_r2q_r3m.err | cut -c1-400; cut -c1-900 gpurun_out/bench_r2q_r3m.json)
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
