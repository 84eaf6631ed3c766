"""Workload for ncu captures of k_block on H1M: first sweep (matrix build) + N steady-state sweeps.
  ncu --set full --import-source on --clock-control none -k regex:k_block --launch-skip 4 --launch-count 2 -o out python tests/tools/prof_h1m.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
from pclean_b200.host_fixture import model as M
from pclean_b200.engine import Engine, load_trace_from_snapshot
from pclean_b200.host_fixture.synth import build_synthetic_hospital
model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(rows, 20260924)
e = Engine(ir, M.InferenceConfig(1, 20))
e.load_observations(obs)
load_trace_from_snapshot(e, ir, model, query.cls, snap)
latent = []
for opt in sys.argv[3:]:
    k, v = opt.split("=")
    if k == "latent": latent = v.split(",")        # then sweep these latent classes once (ncu -k regex:k_latent)
    else: e.set_option(k, int(v))
cls = ir.class_index[query.cls]
for s in range(sweeps):
    st = e.sweep(cls, 1, s + 1)
    print(s, st["total_ms"], e.block_metrics(0)["kernel_ms"], e.block_metrics(1)["kernel_ms"], flush=True)
for name in latent:
    st = e.sweep(ir.class_index[name], 1, sweeps + 1)
    print(name, st["total_ms"], st["kernel_ms"], st["launches"], flush=True)
