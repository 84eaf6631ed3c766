import sys; sys.path.insert(0, '.')  # run from the repo root
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment
from pclean_b200.engine import Engine
cfg = M.InferenceConfig(5, 2, use_mh_instead_of_pg=True)
model, query, dirty, clean, ir, obs = load_experiment("flights")
e = Engine(ir, cfg); e.load_observations(obs)
for opt in sys.argv[1:]:
    k, v = opt.split("="); e.set_option(k, int(v))
try:
    e.init_trace(2); print("init ok", {c: e.table_size(ir.class_index[c]) for c in model.class_order[:-1]})
except Exception as ex:
    print("init failed", ex); sys.exit(0)
tw = ir.class_index["TrackingWebsite"]
keys, refs, cells = e.download_table(tw, 4)
print("TW rows", len(keys), "refs", refs[:12].tolist())
for j in range(6): print("  row", j, [e.decode(cells[v, j]) for v in range(cells.shape[0])])
fl = ir.class_index["Flight"]
keys, refs, cells = e.download_table(fl, 12)
for j in range(3): print("  flight", j, int(refs[j]), [e.decode(cells[v, j]) for v in range(cells.shape[0])])
for it in range(2):
    for name in model.class_order:
        try:
            st = e.sweep(ir.class_index[name], 2, it + 1); print("sweep", it, name, "ok", st["changed_rows"], st["new_rows"])
        except Exception as ex:
            print("sweep", it, name, "FAILED", ex); sys.exit(0)
