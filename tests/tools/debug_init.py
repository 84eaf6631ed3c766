import sys; sys.path.insert(0, '.')  # run from the repo root
import numpy as np
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment
from oracle import Oracle
from pclean_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 245
cfg = M.InferenceConfig(1, 2, use_mh_instead_of_pg=True, rejuv_frequency=10 ** 9)
model, query, dirty, clean, ir, obs = load_experiment(sys.argv[2] if len(sys.argv) > 2 and "=" not in sys.argv[2] else "hospital", max_rows=n)
o = Oracle(ir, cfg, seed=11); o.load_observations(obs); o.initialize_trace(n - 1)
e = Engine(ir, cfg); e.load_observations(obs)
for opt in [a for a in sys.argv[2:] if "=" in a]:
    k, v = opt.split("="); e.set_option(k, int(v))
e.set_option("batch_rows", 1); e.set_option("resample_params", 0); e.set_option("init_rows", n - 1); e.init_trace(11)
cls = ir.class_index[query.cls]
r = n - 1
nb = len(model.classes[query.cls].blocks); ko, wo, so, mo = o.row_move(cls, r, nb)
ke, we, se, me = e.row_move_debug(cls, r, 11, 0, nb)
print("oracle keys", ko.tolist(), wo, so, mo)
print("engine keys", ke.tolist(), we, se, me)
for name in model.class_order[:-1]:
    c = ir.class_index[name]
    print(name, o.table_size(c), e.table_size(c))
