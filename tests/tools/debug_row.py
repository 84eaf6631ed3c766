"""Which engine switch changes a row move?  (H = 4096 synthetic table of tests/test_bench_shape_parity.py)
  python tests/tools/debug_row.py 773 508 710"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.synth import build_synthetic_hospital
from pclean_b200.engine import Engine, load_trace_from_snapshot
from oracle import Oracle

rows = [int(x) for x in sys.argv[1:]] or [773]
n = 50000
cfg = M.InferenceConfig(1, 20)
model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(n, 11, H=4096, P=2048, C=512)
o = Oracle(ir, cfg, seed=11); o.load_observations(obs); o.install_snapshot(ir, model, query.cls, snap); o.begin_sweep()
cls = ir.class_index[query.cls]
for r in rows:
    ko, wo, so, mo = o.clone().row_move(cls, r, 2)
    print("row", r, "oracle w1", wo[1], "sel", so, "ml", mo, "keys", ko[:3].tolist(), flush=True)
    for name, sets in [("default", {}), ("opts=0", {"opts": 0}), ("opts=1", {"opts": 1}), ("opts=2", {"opts": 2}), ("opts=4", {"opts": 4}), ("opts=8", {"opts": 8}),
                       ("memo=0", {"memo": 0}), ("prune=0", {"prune": 0}), ("prune=0 memo=0 opts=0", {"prune": 0, "memo": 0, "opts": 0})]:
        e = Engine(ir, cfg); e.load_observations(obs); load_trace_from_snapshot(e, ir, model, query.cls, snap)
        for k, v in sets.items():
            e.set_option(k, v)
        ke, we, se, me = e.row_move_debug(cls, r, 11, 1, 2)
        print("   ", name, "w1", we[1], "sel", se, "ml", me, "keys", ke[:3].tolist(), "flags", e.download_row_flags(cls, r, r + 1)[0], flush=True)
        e.close()
