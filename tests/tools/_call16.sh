set -x
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_block --launch-skip 4 --launch-count 2 -f -o gpurun_out/prof_kblock_r2n python tests/tools/prof_h1m.py 1000000 3 > gpurun_out/ncu_r2n.log 2>&1; tail -4 gpurun_out/ncu_r2n.log)
(timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_latent --launch-count 5 -f -o gpurun_out/prof_klatent_r2n python tests/tools/prof_h1m.py 1000000 1 latent=Hospital > gpurun_out/ncu_r2n_lat.log 2>&1; tail -4 gpurun_out/ncu_r2n_lat.log)
ncu -i gpurun_out/prof_klatent_r2n.ncu-rep --page raw --csv > gpurun_out/prof_klatent_r2n_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_klatent_r2n.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/prof_klatent_r2n_cudasass.csv.gz
ncu -i gpurun_out/prof_kblock_r2n.ncu-rep --page raw --csv > gpurun_out/prof_kblock_r2n_raw.csv 2>/dev/null
rm -f gpurun_out/prof_klatent_r2n.ncu-rep
ls -la gpurun_out | tail -8
(timeout 600 python scripts/run_h1m_init.py --rows 1000000 > gpurun_out/h1m_init_r2n.json 2> gpurun_out/h1m_init_r2n.err; tail -c 1500 gpurun_out/h1m_init_r2n.json; tail -3 gpurun_out/h1m_init_r2n.err | cut -c1-400)
