"""rents5: engine vs oracle parameter values (keyed prior draws) and row-move weights."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.synth import build_synthetic_rents
from pclean_b200.engine import Engine, load_trace_from_snapshot
from oracle import Oracle

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cfg = M.InferenceConfig(1, K, rejuv_frequency=10 ** 9)
model, query, dirty, truth, ir, obs, snap = build_synthetic_rents(n, 7, n_counties=nc)
print("K", K, "rows", n, "counties", nc, flush=True)
o = Oracle(ir, cfg, seed=7); o.load_observations(obs); o.install_snapshot(ir, model, query.cls, snap); o.begin_sweep()
e = Engine(ir, cfg); e.set_option("param_seed", 7); e.load_observations(obs); load_trace_from_snapshot(e, ir, model, query.cls, snap)
cls = ir.class_index[query.cls]
for r in (0, 5, 100, 1000, 2500)[:int(sys.argv[4]) if len(sys.argv) > 4 else 5]:
    oc = o.clone()
    ko, wo, so, mo = oc.row_move(cls, r, 1)
    ke, we, se, me = e.row_move_debug(cls, r, 7, 1, 1)
    ke2, we2, se2, me2 = e.row_move_debug(cls, r, 7, 2, 1)
    print("   sweep 2: engine", we2[1], we2[-1], "sel", se2, "flags", e.download_row_flags(cls, r, r + 1)[0], "| sweep 1 again:", e.row_move_debug(cls, r, 7, 1, 1)[1][1])
    print("row", r, "oracle", wo[1], "engine", we[1], "keys", ko[1].tolist(), ke[1].tolist(), "state", dirty["State"][r], "br", dirty["Room Type"][r], flush=True)
    # parameters the oracle's move created / read
    bad = 0; seen = 0
    for slot in range(oc.n_slots()):
        vo, _ = oc.param_get(slot)
        if len(vo) == 0:
            continue
        seen += 1
        ve = e.get_param(slot)
        if len(ve) != len(vo) or not np.allclose(vo, ve, rtol=1e-12, atol=0):
            bad += 1
            if bad <= 3:
                print("   slot", slot, ir.slot_key[slot], "oracle", vo[:3], "engine", ve[:3])
    print("   slots with values in the oracle:", seen, "mismatching:", bad, flush=True)
