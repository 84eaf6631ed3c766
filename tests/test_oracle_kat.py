"""Known-answer tests of the oracle against values derived from the reference's formulas
(SURVEY App. E).  The reference ships no golden vectors: these pin the
restatement of add_typos.jl / string_prior.jl / transformed_gaussian.jl / maybe_swap.jl /
time_prior.jl / trace.jl to the formulas, not to reference outputs ("parity unpinned")."""
import json
import math
import os

import numpy as np
import pytest

from pclean_b200 import lowering as LW
from pclean_b200.host_fixture import model as M

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


@pytest.fixture(scope="module")
def oracle(hospital):
    from oracle import Oracle
    model, query, dirty, clean, ir, obs = hospital
    return Oracle(ir, M.InferenceConfig(1, 2), seed=0)


def test_addtypos_kat(oracle):
    for obs, word, max_typos, want in GOLD["addtypos"]:
        got = oracle.addtypos(obs, word, max_typos)
        assert got == pytest.approx(want, rel=1e-9, abs=1e-9), (obs, word)


def test_osa_vs_true_damerau(oracle):
    assert oracle.edit_distance("ca", "abc") == 3          # optimal string alignment
    assert oracle.edit_distance("brimingham", "birmingham") == 1
    assert oracle.edit_distance("", "abc") == 3
    oracle.L.oracle_set_true_damerau(oracle.h, 1)
    assert oracle.edit_distance("ca", "abc") == 2          # unrestricted Damerau-Levenshtein
    oracle.L.oracle_set_true_damerau(oracle.h, 0)


def test_stringprior_kat(oracle):
    for s, lo, hi, want in GOLD["stringprior"]:
        assert oracle.stringprior(s, lo, hi) == pytest.approx(want, rel=1e-9)


def _val(tag, i=0, d=0.0):
    return LW.Value(tag, i, d)


def test_other_densities(oracle):
    L = oracle.L
    # TransformedGaussian: args (mean, std, transformation)
    ir = oracle.ir
    def logd(dist, obs, args):
        arr = (LW.Value * len(args))(*args)
        return L.oracle_logdensity(oracle.h, dist, obs, len(args), arr)
    # crp
    assert L.oracle_crp_logprior(22, 0.0, 1.0, 999) == pytest.approx(-3.8167128256, rel=1e-9)
    # time prior atom
    assert logd(M.DIST_TIME_PRIOR, _val(LW.VAL_STR, 0), [_val(LW.VAL_LIST, 0)]) == pytest.approx(-7.2723983926, rel=1e-9)


def test_logsumexp(oracle):
    x = (LW.C.c_double * 3)(-1.0, -2.0, -3.0)
    assert oracle.L.oracle_logsumexp(3, x) == pytest.approx(math.log(sum(math.exp(v) for v in (-1, -2, -3))), rel=1e-12)
    assert oracle.L.oracle_logsumexp(0, x) == -math.inf
