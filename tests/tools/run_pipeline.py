"""Engine-only pipeline on a shipped benchmark: initialize_trace + run_inference! on the GPU, F1
against the clean table; optionally the oracle (CPU restatement) beside it for the same config."""
import sys, time, json; sys.path.insert(0, '.')  # run from the repo root
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment
from pclean_b200.host_fixture.analysis import evaluate_accuracy
from pclean_b200.engine import Engine

name = sys.argv[1]
max_rows = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
with_oracle = len(sys.argv) > 3 and sys.argv[3] == "oracle"
pg_particles = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # > 0: particle Gibbs with this many particles (BASELINE configs[1], [2]) instead of the shipped MH
CFG = {"hospital": M.InferenceConfig(1, 2, use_mh_instead_of_pg=True),
       "rents": M.InferenceConfig(1, 2, use_mh_instead_of_pg=True, rejuv_frequency=500),
       "flights": M.InferenceConfig(5, 2, use_mh_instead_of_pg=True)}[name]
if pg_particles:
    CFG = M.InferenceConfig(CFG.num_iters, pg_particles, rejuv_frequency=CFG.rejuv_frequency)
model, query, dirty, clean, ir, obs = load_experiment(name, max_rows=max_rows)
n = obs.n_rows
cls = ir.class_index[query.cls]
cols = list(query.cleanmap.keys()); verts = [query.cleanmap[c] - 1 for c in cols]
out = {"benchmark": name, "rows": n, "config": "InferenceConfig(%d, %d; use_mh_instead_of_pg=%s, rejuv_frequency=%d)" % (CFG.num_iters, CFG.num_particles, CFG.use_mh_instead_of_pg, CFG.rejuv_frequency)}
t0 = time.time(); e = Engine(ir, CFG); e.load_observations(obs)
for opt in sys.argv[5:]:
    k, v = opt.split("="); e.set_option(k, int(v)); out.setdefault("options", {})[k] = int(v)
t1 = time.time(); e.init_trace(1); t2 = time.time()
def f1e():
    cells = e.download_cells(cls, verts, n)
    return evaluate_accuracy(dirty, clean, {c: [e.decode(cells[k, r]) for r in range(n)] for k, c in enumerate(cols)}, cols)
a0 = f1e(); t3 = time.time(); st = e.run_inference(1); t4 = time.time(); a1 = f1e()
out["engine"] = {"init_s": t2 - t1, "f1_after_init": a0["f1"], "sweeps_s": t4 - t3, "f1": a1["f1"], "precision": a1["precision"], "recall": a1["recall"],
                 "tables": {c: e.table_size(ir.class_index[c]) for c in model.class_order[:-1]}, "dummy_draws": st["dummy_draws"]}
if with_oracle:
    from oracle import Oracle
    o = Oracle(ir, CFG, seed=1); o.load_observations(obs)
    t0 = time.time(); o.initialize_trace(); t1 = time.time()
    def f1o():
        cells = o.get_cells(cls, verts)
        return evaluate_accuracy(dirty, clean, {c: [o.decode(cells[k, r]) for r in range(n)] for k, c in enumerate(cols)}, cols)
    b0 = f1o(); t2 = time.time(); o.run_inference(); t3 = time.time(); b1 = f1o()
    out["oracle_cpu"] = {"init_s": t1 - t0, "f1_after_init": b0["f1"], "sweeps_s": t3 - t2, "f1": b1["f1"]}
print(json.dumps(out))
