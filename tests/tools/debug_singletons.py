"""Singleton-hospital rows of the H = 4096 synthetic table: oracle vs engine weight of a new-row
particle, next to the reference counts along the row's chain (which cascade depth disagrees?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.synth import build_synthetic_hospital
from pclean_b200.engine import Engine, load_trace_from_snapshot
from oracle import Oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
cfg = M.InferenceConfig(1, 4)
model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(n, 11, H=4096, P=2048, C=512)
o = Oracle(ir, cfg, seed=11); o.load_observations(obs); o.install_snapshot(ir, model, query.cls, snap); o.begin_sweep()
e = Engine(ir, cfg); e.load_observations(obs); load_trace_from_snapshot(e, ir, model, query.cls, snap)
cls = ir.class_index[query.cls]
t = truth
hcount = np.bincount(t["row_h"], minlength=len(t["name"]))
pcount = np.bincount(t["h_place"], minlength=len(t["city"]))
ccount = np.bincount(t["place_county"], minlength=len(t["county_name"]))
tcount = np.bincount(t["h_type"], minlength=len(t["types"]))
rows = [r for r in range(0, 4096) if hcount[t["row_h"][r]] == 1][:int(sys.argv[2]) if len(sys.argv) > 2 else 120]
for r in rows:
    h = t["row_h"][r]; p = t["h_place"][h]; c = t["place_county"][p]
    ko, wo, so, mo = o.clone().row_move(cls, r, 2)
    ke, we, se, me = e.row_move_debug(cls, r, 11, 1, 2)
    print(r, "diff %.6f" % (we[1] - wo[1]), "w_oracle %.4f" % wo[1], "place_cnt", pcount[p], "county_cnt", ccount[c], "type_cnt", tcount[t["h_type"][h]],
          "keys_o", ko[1].tolist(), "keys_e", ke[1].tolist(), flush=True)
