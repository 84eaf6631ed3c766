set -x
free -g | head -2
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2w_h1m_8gpu.json 2> gpurun_out/bench_r2w_h1m_8gpu.err; tail -2 gpurun_out/bench_r2w_h1m_8gpu.err | cut -c1-300; cut -c1-330 gpurun_out/bench_r2w_h1m_8gpu.json)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload r10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2w_r10m_8gpu.json 2> gpurun_out/bench_r2w_r10m_8gpu.err; tail -2 gpurun_out/bench_r2w_r10m_8gpu.err | cut -c1-300; cut -c1-330 gpurun_out/bench_r2w_r10m_8gpu.json)
