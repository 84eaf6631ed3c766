"""The on-disk IR (PCLIRv1, pclean_b200/irfile.py): what a Julia host would write
(julia/PCleanB200.jl `write_ir`) and what `pclean_load_model_file` reads."""
import os

import numpy as np
import pytest

from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment
from pclean_b200.irfile import load_ir, save_ir
from pclean_b200.lowering import ModelIR


@pytest.mark.parametrize("name,rows", [("hospital", 200), ("rents", 500), ("flights", 300)])
def test_ir_file_round_trip(tmp_path, name, rows):
    model, query, dirty, clean, ir, obs = load_experiment(name, max_rows=rows)
    path = os.path.join(tmp_path, f"{name}.pclir")
    save_ir(path, ir, obs)
    back = load_ir(path)
    a, b = ir.as_ctypes(), back.as_ctypes()
    for fname, ctype in ModelIR._fields_:
        if fname in ir._arrays:
            assert np.array_equal(np.asarray(ir._arrays[fname]).reshape(-1), back.entries[fname]), fname
        else:
            assert getattr(a, fname) == getattr(b, fname), fname
    assert back.strings == ir.strings and back.class_names == list(model.class_order)
    o2 = back.observations()
    assert o2.cls == obs.cls and o2.n_rows == obs.n_rows and o2.n_cols == obs.n_cols
    assert np.array_equal(o2._keep[0], obs._keep[0]) and np.array_equal(o2._keep[1].reshape(-1), np.ascontiguousarray(obs._keep[1]).reshape(-1))


@pytest.mark.gpu
def test_engine_from_ir_file_equals_engine_from_memory(tmp_path):
    """model + observations read by the library from the file (pclean_load_model_file /
    pclean_load_observations_file) give the same initialize_trace + run_inference! as the in-memory IR"""
    from pclean_b200.engine import Engine
    cfg = M.InferenceConfig(1, 2, use_mh_instead_of_pg=True)
    model, query, dirty, clean, ir, obs = load_experiment("hospital", max_rows=400)
    path = os.path.join(tmp_path, "hospital.pclir")
    save_ir(path, ir, obs)
    cls = ir.class_index[query.cls]
    verts = sorted(v - 1 for v in query.cleanmap.values())
    res = []
    for from_file in (False, True):
        if from_file:
            e = Engine(path, cfg)
            e.load_observations_file(path)
        else:
            e = Engine(ir, cfg)
            e.load_observations(obs)
        e.init_trace(5)
        st = e.run_inference(5)
        cells = e.download_cells(cls, verts, 400)
        res.append(([[e.decode(cells[k, r]) for r in range(400)] for k in range(len(verts))], e.download_logweights(cls, 400), st["changed_rows"]))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
