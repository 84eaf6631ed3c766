"""The Plan -> stars lowering (pclean_b200/csrc/lower.hpp, host C++) on the three shipped programs:
block shapes that the CUDA kernels rely on, checked without a GPU.  Counts follow SURVEY App. A /
DESIGN.md section 3."""
import ctypes as C
import os
import subprocess

import pytest

from pclean_b200.host_fixture.experiments import load_experiment

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lower():
    so = os.path.join(HERE, "_lower_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                           "-I", "/usr/local/cuda/include", "-o", so, os.path.join(HERE, "lower_host.cpp")])
    return C.CDLL(so)


def summary(L, ir, obs, cls, block, latent=False, own=(), drop=()):
    cols = [obs.vertex_of_col[i] for i in range(obs.n_cols) if obs.vertex_of_col[i] not in drop]
    arr = (C.c_int * max(1, len(cols)))(*cols)
    own_arr = (C.c_int * max(1, len(own)))(*own)
    out = (C.c_int * 9)()
    err = C.create_string_buffer(512)
    cir = ir.as_ctypes()
    rc = L.lower_summary(C.byref(cir), ir.class_index[cls], block, int(latent), ir.class_index[ir_obs_class(ir, obs)],
                         arr, len(cols), own_arr, len(own), out, err, 512)
    keys = ["ok", "root", "stars", "terms", "roots", "rootless", "root_terms", "sampled", "fillins"]
    return dict(zip(keys, list(out)), error=err.value.decode() if rc else None)


def ir_obs_class(ir, obs):
    return [k for k, v in ir.class_index.items() if v == obs.cls][0]


def test_hospital_blocks(lower):
    model, query, dirty, clean, ir, obs = load_experiment("hospital")
    b0 = summary(lower, ir, obs, "Record", 0)
    b1 = summary(lower, ir, obs, "Record", 1)
    assert (b0["stars"], b0["terms"]) == (15, 28) and (b1["stars"], b1["terms"]) == (5, 9)
    hosp = summary(lower, ir, obs, "Hospital", 0, latent=True)
    assert hosp["ok"] and hosp["roots"] == 9 and hosp["stars"] == 14 and hosp["terms"] == 20
    place = summary(lower, ir, obs, "Place", 0, latent=True)
    assert place["roots"] == 2 and place["stars"] == 4


def test_rents_patterns(lower):
    model, query, dirty, clean, ir, obs = load_experiment("rents", max_rows=2000)
    full = summary(lower, ir, obs, "Obs", 0)
    assert full["ok"] and full["stars"] >= 1 and full["rootless"] == 0
    state_v = [v for v in (obs.vertex_of_col[i] for i in range(obs.n_cols)) if v == 9]
    missing_state = summary(lower, ir, obs, "Obs", 0, drop=set(state_v))
    assert missing_state["ok"] and missing_state["stars"] == full["stars"] + 1          # the state becomes a choice star of the new-row branch
    # County rows: countykey observed only -> name and state sites; state observed too -> name site only
    both = summary(lower, ir, obs, "County", 0, latent=True, own=[1])
    name_only = summary(lower, ir, obs, "County", 0, latent=True, own=[1, 7])
    assert both["roots"] == 2 and both["terms"] == 2 and name_only["roots"] == 1 and name_only["terms"] == 1
    none = summary(lower, ir, obs, "County", 0, latent=True)
    assert not none["ok"]                                                              # an unobserved Unmodeled key cannot be proposed


def test_flights_blocks(lower):
    model, query, dirty, clean, ir, obs = load_experiment("flights")
    b0 = summary(lower, ir, obs, "Obs", 0)
    b1 = summary(lower, ir, obs, "Obs", 1)
    b2 = summary(lower, ir, obs, "Obs", 2)
    assert b0["ok"] and b0["stars"] == 1 and b0["fillins"] == 4           # a new Flight draws its four times from the TimePrior proposal
    assert b1["ok"] and b1["stars"] == 1 and b1["terms"] == 1             # equality with the observed website name
    assert b2["ok"] and b2["rootless"] == 1 and b2["root_terms"] == 4     # no enumeration: four MaybeSwap likelihoods
    b2m = summary(lower, ir, obs, "Obs", 2, drop={21, 25})
    assert b2m["root_terms"] == 2 and b2m["sampled"] == 2                 # absent observations are sampled with random()
    f1 = summary(lower, ir, obs, "Flight", 1, latent=True, own=[3])
    assert f1["ok"] and f1["roots"] == 4 and f1["terms"] == 4             # four TimePrior sites with External MaybeSwap terms
    f0 = summary(lower, ir, obs, "Flight", 0, latent=True, own=[3])
    assert not f0["ok"] and "enumerat" in f0["error"]                     # block 1 of Flight enumerates nothing
