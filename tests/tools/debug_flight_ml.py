"""Flight latent rows: the block log-marginal the oracle and the engine report, key by key
(tests/test_engine_parity.py::test_flights_latent_flight_parity compares cells + selection)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pclean_b200.host_fixture import model as M
import test_engine_parity as T
cfg = M.InferenceConfig(1, 20)
model, query, ir, dirty, clean, obs, o, e = T._setup_flights(cfg)
cls = ir.class_index["Flight"]
cm = model.classes["Flight"]
n_normal = sum(1 for n in cm.nodes if not isinstance(n, M.ExternalLikelihoodNode))
keys, _ = o.table_keys(cls)
nd = 0
for key in keys[:: max(1, len(keys) // 40)]:
    oc = o.clone()
    ko, wo, so, mo = oc.row_move(cls, int(key), len(cm.blocks))
    cells_e, se, me = e.latent_move_debug(cls, int(key), 3, 2, n_normal)
    d = mo - me
    nd += abs(d) > 1e-9 * max(1.0, abs(mo))
    print(int(key), so, se, "ml oracle %.12g engine %.12g diff %.6g" % (mo, me, d), "w", np.asarray(wo)[:4].tolist())
print("differing", nd, "blocks", len(cm.blocks))
