import sys; sys.path.insert(0, '.')  # run from the repo root
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment
from oracle import Oracle, export_snapshot
from pclean_b200.engine import Engine, load_trace_from_snapshot
cfg = M.InferenceConfig(1, 2, use_mh_instead_of_pg=True, rejuv_frequency=500)
n = 3000
model, query, dirty, clean, ir, obs = load_experiment("rents", max_rows=n)
o = Oracle(ir, cfg, seed=4); o.load_observations(obs); o.initialize_trace()
snap = export_snapshot(o, ir, model, query.cls)
cls = ir.class_index[query.cls]
cols = list(query.cleanmap.keys()); verts = [query.cleanmap[c] - 1 for c in cols]
e = Engine(ir, cfg); e.load_observations(obs); load_trace_from_snapshot(e, ir, model, query.cls, snap)
ce = e.download_cells(cls, verts, n); co = o.get_cells(cls, verts)
for k, c in enumerate(cols):
    diff = [(r, o.decode(co[k, r]), e.decode(ce[k, r])) for r in range(n) if o.decode(co[k, r]) != e.decode(ce[k, r])]
    print(c, verts[k], "mismatches", len(diff), diff[:5])
