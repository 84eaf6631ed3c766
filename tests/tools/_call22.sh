set -x
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r2_bench.log 2>&1; wc -l gpurun_out/launches_r2.csv)
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_block --launch-skip 4 --launch-count 2 -f -o gpurun_out/prof_kblock_r2f python tests/tools/prof_h1m.py 1000000 3 > gpurun_out/ncu_r2f.log 2>&1; tail -3 gpurun_out/ncu_r2f.log)
ncu -i gpurun_out/prof_kblock_r2f.ncu-rep --page raw --csv > gpurun_out/prof_kblock_r2f_raw.csv 2>/dev/null
(timeout 900 ncu --set full --clock-control none -k regex:k_latent --launch-count 5 -f -o gpurun_out/prof_klatent_r2f python tests/tools/prof_h1m.py 1000000 1 latent=Hospital > gpurun_out/ncu_r2f_lat.log 2>&1; tail -2 gpurun_out/ncu_r2f_lat.log)
ncu -i gpurun_out/prof_klatent_r2f.ncu-rep --page raw --csv > gpurun_out/prof_klatent_r2f_raw.csv 2>/dev/null; rm -f gpurun_out/prof_klatent_r2f.ncu-rep
(timeout 400 python bench.py --sweep all --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f_h1m_all.json 2> gpurun_out/bench_r2f_h1m_all.err; cut -c1-300 gpurun_out/bench_r2f_h1m_all.json)
(timeout 300 python bench.py --workload rents --steps 10 --warmup 3 > gpurun_out/bench_r2f_rents.json 2> gpurun_out/bench_r2f_rents.err; cut -c1-300 gpurun_out/bench_r2f_rents.json)
(timeout 300 python bench.py --workload rents --sweep all --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f_rents_all.json 2> gpurun_out/bench_r2f_rents_all.err; cut -c1-300 gpurun_out/bench_r2f_rents_all.json)
(timeout 300 python bench.py --workload flights --steps 10 --warmup 3 > gpurun_out/bench_r2f_flights.json 2> gpurun_out/bench_r2f_flights.err; cut -c1-300 gpurun_out/bench_r2f_flights.json)
(timeout 300 python bench.py --workload flights --sweep all --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f_flights_all.json 2> gpurun_out/bench_r2f_flights_all.err; cut -c1-300 gpurun_out/bench_r2f_flights_all.json)
