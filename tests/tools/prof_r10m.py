"""Workload for ncu captures of k_block on the rents-schema synthetic table (bench.py --workload r10m):
first sweep (matrix build) + N steady-state sweeps at K = 50.
  ncu --set full --import-source on --clock-control none -k regex:k_block --launch-skip 2 --launch-count 1 -o out python tests/tools/prof_r10m.py 1000000 3"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
from pclean_b200.host_fixture import model as M
from pclean_b200.engine import Engine, load_trace_from_snapshot
from pclean_b200.host_fixture.synth import build_synthetic_rents
model, query, dirty, truth, ir, obs, snap = build_synthetic_rents(rows, 20260925)
e = Engine(ir, M.InferenceConfig(1, 50, rejuv_frequency=500))
e.load_observations(obs)
load_trace_from_snapshot(e, ir, model, query.cls, snap)
for opt in sys.argv[3:]:
    k, v = opt.split("="); e.set_option(k, int(v))
cls = ir.class_index[query.cls]
for s in range(sweeps):
    st = e.sweep(cls, 1, s + 1)
    print(s, st["total_ms"], st["kernel_ms"], e.block_metrics(0)["kernel_ms"], flush=True)
