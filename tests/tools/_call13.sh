set -x
(timeout 900 python -m pytest tests/test_engine_parity.py -q -m gpu -x -s -k "compaction or sequential_sweep or pipeline_hospital or full_engine" > gpurun_out/t_r2l.log 2>&1; tail -12 gpurun_out/t_r2l.log | cut -c1-1500)
(timeout 400 python bench.py --sweep all --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2l_all.json 2> gpurun_out/bench_r2l_all.err; tail -3 gpurun_out/bench_r2l_all.err | cut -c1-600; python -c "import json; d=json.load(open(\"gpurun_out/bench_r2l_all.json\")); print(d[\"ms_per_step\"], d[\"value\"], d[\"e2e\"][\"value\"], d[\"sweep\"])")
(timeout 300 python tests/tools/debug_flight_ml.py > gpurun_out/dbg_flight_ml.log 2>&1; tail -8 gpurun_out/dbg_flight_ml.log | cut -c1-400)
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_block --launch-skip 4 --launch-count 2 -f -o gpurun_out/prof_kblock_r2l python tests/tools/prof_h1m.py 1000000 3 > gpurun_out/ncu_r2l.log 2>&1; tail -4 gpurun_out/ncu_r2l.log)
(timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_latent --launch-count 5 -f -o gpurun_out/prof_klatent_r2l python tests/tools/prof_h1m.py 1000000 1 latent=Hospital > gpurun_out/ncu_r2l_lat.log 2>&1; tail -4 gpurun_out/ncu_r2l_lat.log)
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_dp_matrix --launch-skip 6 --launch-count 3 -f -o gpurun_out/prof_kdp_r2l python tests/tools/prof_h1m.py 1000000 1 > gpurun_out/ncu_r2l_dp.log 2>&1; tail -4 gpurun_out/ncu_r2l_dp.log)
for n in prof_klatent_r2l prof_kdp_r2l; do ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null; ncu -i gpurun_out/$n.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${n}_source.csv.gz; done
ncu -i gpurun_out/prof_kblock_r2l.ncu-rep --page raw --csv > gpurun_out/prof_kblock_r2l_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
rm -f gpurun_out/prof_kdp_r2l.ncu-rep; [ $(stat -c %s gpurun_out/prof_klatent_r2l.ncu-rep) -gt 25000000 ] && rm -f gpurun_out/prof_klatent_r2l.ncu-rep
rm -f gpurun_out/prof_kblock_r1*.ncu-rep
