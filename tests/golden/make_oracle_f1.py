"""Regenerates tests/golden/oracle_f1.json: the oracle's end-to-end accuracy on the three
shipped datasets with the shipped configs (experiments/*/run.jl), seed 1.  These numbers are
outputs of OUR restatement (the reference cannot be executed here), kept as a regression pin."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.analysis import evaluate_accuracy
from pclean_b200.host_fixture.experiments import load_experiment
from oracle import Oracle

CFG = {"hospital": M.InferenceConfig(1, 2, use_mh_instead_of_pg=True),
       "rents": M.InferenceConfig(1, 2, use_mh_instead_of_pg=True, rejuv_frequency=500),
       "flights": M.InferenceConfig(5, 2, use_mh_instead_of_pg=True)}


def run(name, seed=1):
    model, query, dirty, clean, ir, obs = load_experiment(name)
    o = Oracle(ir, CFG[name], seed=seed)
    o.load_observations(obs)
    o.initialize_trace()
    o.run_inference()
    cls = ir.class_index[query.cls]
    cols = list(query.cleanmap.keys())
    cells = o.get_cells(cls, [query.cleanmap[c] - 1 for c in cols])
    ours = {c: [o.decode(cells[k, r]) for r in range(cells.shape[1])] for k, c in enumerate(cols)}
    return evaluate_accuracy(dirty, clean, ours, cols)


if __name__ == "__main__":
    out = {n: run(n) for n in ("hospital", "flights", "rents")}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_f1.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
