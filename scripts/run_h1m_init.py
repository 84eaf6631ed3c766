"""Engine-only run on the synthetic hospital-schema table of bench.py (H1M): initialize_trace on the
device, then full sweeps over every class; accuracy against the generator's clean table."""
import sys, time, json, argparse; sys.path.insert(0, '.')
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.synth import build_synthetic_hospital
from pclean_b200.host_fixture.analysis import evaluate_accuracy
from pclean_b200.engine import Engine

ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=1_000_000); ap.add_argument("--particles", type=int, default=20)
ap.add_argument("--sweeps", type=int, default=1); ap.add_argument("--table-cap", type=int, default=12288); ap.add_argument("--seed", type=int, default=20260924); a = ap.parse_args()
scale = {} if a.rows >= 1_000_000 else dict(H=max(64, a.rows // 256), P=max(32, a.rows // 512), C=max(16, a.rows // 2048))
model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(a.rows, a.seed, **scale)
cfg = M.InferenceConfig(a.sweeps, a.particles)
e = Engine(ir, cfg); e.load_observations(obs); e.set_option("table_cap", a.table_cap)
if a.rows >= 1_000_000:      # reserve per class what the generator can produce (+ duplicates of batched initialisation)
    for cname, rows in (("Measure", 4096), ("Condition", 1024), ("HospitalType", 1024), ("County", 8192), ("Place", 16384), ("Hospital", 2 * a.table_cap)):
        e.reserve_table(ir.class_index[cname], rows)
t0 = time.time(); e.init_trace(a.seed); t1 = time.time()
cls = ir.class_index[query.cls]
cols = list(query.cleanmap.keys()); verts = [query.cleanmap[c] - 1 for c in cols]
n = a.rows
def acc(sample=200000):
    m = min(n, sample)
    cells = e.download_cells(cls, verts, n)
    ours = {c: [e.decode(cells[k, r]) for r in range(m)] for k, c in enumerate(cols)}
    d = {c: dirty[c][:m] for c in cols}; t = {c: truth["clean"][c][:m] for c in cols}
    return evaluate_accuracy(d, t, ours, cols)
out = {"rows": n, "particles": a.particles, "init_s": t1 - t0, "init_rows_per_s": n / (t1 - t0),
       "tables_after_init": {c: e.table_size(ir.class_index[c]) for c in model.class_order[:-1]}}
a0 = acc(); out["f1_after_init"] = a0["f1"]
print(json.dumps(out), flush=True)          # the initialisation alone (the sweeps follow)
t2 = time.time(); st = e.run_inference(a.seed); t3 = time.time()
a1 = acc(); out.update({"sweeps": a.sweeps, "sweeps_s": t3 - t2, "f1": a1["f1"], "precision": a1["precision"], "recall": a1["recall"],
                        "matrices_gib": e.matrix_bytes() / 2**30, "new_rows": st["new_rows"], "changed_rows": st["changed_rows"]})
print(json.dumps(out))
