"""GPU parity on the shape bench.py measures (synthetic hospital-schema, H = 4096 hospitals):
the multi-stride pruning loops, the star-marginal memo switched off and on, a whole synchronous
sweep against the oracle moving the same rows on the same frozen snapshot, and the parameter /
Pitman-Yor moves against the oracle's under the shared keyed RNG (include/pclean_rng.h).

The oracle is the builder's CPU restatement of the reference (oracle/pclean_oracle.cpp); the
reference itself cannot run here, so "parity" below means engine == oracle."""
import numpy as np
import pytest

from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _setup_synth(config, n_rows, H=4096, seed=11, **kw):
    from oracle import Oracle
    from pclean_b200.engine import Engine, load_trace_from_snapshot
    from pclean_b200.host_fixture.synth import build_synthetic_hospital
    model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(n_rows, seed, H=H, P=H // 2, C=H // 8, **kw)
    o = Oracle(ir, config, seed=seed)
    o.load_observations(obs)
    o.install_snapshot(ir, model, query.cls, snap)
    o.begin_sweep()                       # sweep index 1
    e = Engine(ir, config)
    e.set_option("param_seed", seed)      # the trace carries no parameter values: both sides draw them from the same keyed stream
    e.load_observations(obs)
    load_trace_from_snapshot(e, ir, model, query.cls, snap)
    return model, query, ir, dirty, truth, o, e


def _typo_rows(dirty, truth, lo, hi, limit):
    clean = truth["clean"]
    out = []
    for r in range(lo, hi):
        if any(dirty[c][r] != clean[c][r] for c in dirty):
            out.append(r)
            if len(out) >= limit:
                break
    return out


def test_row_move_parity_h4096_multistride():
    """H = 4096 candidates (the benchmark's table size): the 16-wide pruning loops run 8 strides per
    row; prune = 1 (integer bound) and prune = 0 (exhaustive) both reproduce the oracle's run_smc!
    on singleton hospitals (garbage-collection cascade + new-row branch), ordinary rows and rows
    with typos."""
    cfg = M.InferenceConfig(1, 20)
    n = 50000
    model, query, ir, dirty, truth, o, e = _setup_synth(cfg, n)
    cls = ir.class_index[query.cls]
    nb = 2
    # rows 0..4095 reference every hospital once (the tail of the Zipf law has no other reference)
    rows = list(range(5, 4096, 64)) + list(range(4100, 4100 + 40)) + _typo_rows(dirty, truth, 8192, n, 100)
    assert len(rows) >= 200
    want = {}
    for r in rows:
        want[r] = o.clone().row_move(cls, int(r), nb)
    for prune in (1, 0):
        e.set_option("prune", prune)
        bad = []
        for r in rows:
            ko, wo, so, mo = want[r]
            ke, we, se, me = e.row_move_debug(cls, int(r), 11, 1, nb)
            dummy = bool(e.download_row_flags(cls, int(r), int(r) + 1)[0] & 1)     # reported, compared like any other row
            ok = so == se and np.allclose(wo, we, rtol=RTOL, atol=1e-9) and np.isclose(mo, me, rtol=RTOL, atol=1e-9)
            ok = ok and (ko[1:] == ke[1:]).all()
            if not ok:
                kd = [(k, ko[k].tolist(), ke[k].tolist()) for k in range(1, ko.shape[0]) if (ko[k] != ke[k]).any()]
                bad.append(dict(prune=prune, row=r, dummy=dummy, sel=(so, se), log_ml=(mo, me), max_w_diff=float(np.max(np.abs(wo - we))), key_diffs=kd[:3]))
        assert not bad, (len(bad), bad[:3])


def test_memo_off_and_on_give_the_same_sweep():
    """the star-marginal memo (device hash table shared by the rows of a launch) must be invisible:
    a full observation sweep with memo = 0 and with memo = 1 selects the same rows with the same
    log-weights (rows that share a key read a value another row computed from the same inputs)"""
    cfg = M.InferenceConfig(1, 20)
    n = 30000
    res = []
    for memo in (0, 1):
        model, query, ir, dirty, truth, o, e = _setup_synth(cfg, n)
        cls = ir.class_index[query.cls]
        e.set_option("memo", memo)
        sts = [e.sweep(cls, 5, 1), e.sweep(cls, 5, 2)]          # second sweep: entries persisted for choice stars are reused
        rec = model.classes[query.cls]
        res.append((e.download_assignment(cls, rec.names["hosp"] - 1, n), e.download_assignment(cls, rec.names["metric"] - 1, n),
                    e.download_logweights(cls, n), sts))
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all()
    assert np.allclose(res[0][2], res[1][2], rtol=1e-12, atol=0)
    assert res[0][3][0]["changed_rows"] == res[1][3][0]["changed_rows"] and res[0][3][0]["new_rows"] == res[1][3][0]["new_rows"]


def test_synchronous_sweep_matches_oracle_on_frozen_snapshot():
    """One synchronous sweep of 20,000 synthetic rows (all rows moved against the same table
    snapshot, the default mode bench.py times): for a sample of rows, the row the engine installed
    and its log-weight equal what the oracle's run_smc! gives for that row on a clone of the same
    snapshot (clone per row = frozen snapshot)."""
    cfg = M.InferenceConfig(1, 20)
    n = 20000
    model, query, ir, dirty, truth, o, e = _setup_synth(cfg, n)
    cls = ir.class_index[query.cls]
    rec = model.classes[query.cls]
    fks = [rec.names["hosp"] - 1, rec.names["metric"] - 1]
    e.set_option("resample_params", 0)
    before = [e.download_assignment(cls, f, n) for f in fks]
    existing = [set(o.table_keys(ir.class_index[t])[0].tolist()) for t in ("Hospital", "Measure")]
    st = e.sweep(cls, 11, 1)
    after = [e.download_assignment(cls, f, n) for f in fks]
    lw = e.download_logweights(cls, n)
    flags = e.download_row_flags(cls, 0, n)
    rows = list(range(3, 4096, 101)) + list(range(4096, n, 211)) + _typo_rows(dirty, truth, 8192, n, 60)
    # and every row where some particle drew the dummy of a StringPrior choice: the reference replaces it by
    # random(StringPrior) (block_proposal.jl:58-60) and scores the observation against that string; the engine draws
    # the same string from the same keyed stream, scores it with an inline DP and interns it if its particle is applied
    rows = sorted(set(rows + [int(r) for r in np.nonzero(flags & 1)[0][:40]]))
    assert len(rows) >= 150
    bad = []
    n_dummy = 0
    for r in rows:
        ko, wo, so, mo = o.clone().row_move(cls, int(r), 2)
        ok = np.isclose(mo, lw[r], rtol=RTOL, atol=1e-9)
        if flags[r] & 1:
            n_dummy += 1      # a particle drew a StringPrior dummy: compared like any other row (random(StringPrior) is drawn on the device)
        for b in range(2):
            if so == 0:
                ok = ok and after[b][r] == before[b][r]            # the retained particle: nothing changes
            elif int(ko[so, b]) in existing[b]:
                ok = ok and after[b][r] == ko[so, b]
            else:
                ok = ok and after[b][r] not in existing[b]         # a row created by this move
        if not ok:
            bad.append((r, so, ko[so].tolist(), [int(after[0][r]), int(after[1][r])], mo, float(lw[r])))
    assert st["rows"] == n and not bad and st["dummy_draws"] > 0 and n_dummy > 0, (len(bad), n_dummy, bad[:3], st)


def _param_parity(name, cfg, classes, seed, max_rows=None, mean_rtol=1e-9, init_cfg=None):
    from oracle import Oracle, export_snapshot
    from pclean_b200.engine import Engine, load_trace_from_snapshot
    model, query, dirty, clean, ir, obs = load_experiment(name, max_rows=max_rows)
    o = Oracle(ir, init_cfg or cfg, seed=seed)
    o.load_observations(obs)
    o.initialize_trace()
    o.set_config(cfg)
    snap = export_snapshot(o, ir, model, query.cls)
    e = Engine(ir, cfg)
    e.load_observations(obs)
    load_trace_from_snapshot(e, ir, model, query.cls, snap)
    o.set_epochs(0)                        # a fresh engine starts every resampling counter at 0
    bad = []
    checked = {"slots": 0, "py": 0}
    for cname in classes:
        c = ir.class_index[cname]
        o.resample_class(c)
        e.sweep(c, seed, 1)                # resamples the class's parameters + PY hyper-parameters, then moves its rows
        if cname != query.cls:
            so, do, _ = o.get_py(c)
            se, de = e.get_py_params(c)
            checked["py"] += 1
            if not (np.isclose(so, se, rtol=1e-12, atol=0) and np.isclose(do, de, rtol=1e-12, atol=0)):
                bad.append(("py", cname, so, do, se, de))
    for slot in range(o.n_slots()):
        vo, _ = o.param_get(slot)
        if len(vo) == 0:
            continue
        ve = e.get_param(slot)
        checked["slots"] += 1
        if len(ve) != len(vo) or not np.allclose(vo, ve, rtol=mean_rtol, atol=1e-12):
            bad.append(("slot", slot, vo[:4].tolist(), ve[:4].tolist()))
    return bad, checked


def test_dirichlet_and_pitman_yor_moves_match_oracle_hospital():
    """ProportionsParameter Gibbs steps (choose_proportionally.jl:70-74: Dirichlet(alpha + counts), the
    counts taken from the live rows on the device) and the Pitman-Yor hyper-parameter MH
    (trace.jl:83-107) of every hospital latent class, against the oracle under the same keyed streams"""
    cfg = M.InferenceConfig(1, 2, use_mh_instead_of_pg=True)
    bad, checked = _param_parity("hospital", cfg, ["County", "Place", "Condition", "Measure", "HospitalType", "Hospital"], seed=9)
    assert checked["py"] == 6 and checked["slots"] >= 3 and not bad, (checked, bad[:4])


def test_beta_moves_match_oracle_flights():
    """ProbParameter Gibbs steps (maybe_swap.jl:87-89: Beta(a + differing, b + equal) per tracking
    website, counts taken on the device) of the flights observation class, against the oracle"""
    cfg = M.InferenceConfig(1, 2, use_mh_instead_of_pg=True)
    bad, checked = _param_parity("flights", cfg, ["Obs"], seed=4)
    assert checked["slots"] >= 20 and not bad, (checked, bad[:4])


def test_mean_parameter_moves_match_oracle_rents():
    """MeanParameter Gibbs steps (add_noise.jl:74-82) of the rents observation class: the engine sums
    the moments on the device in row order, the oracle accumulates them incrementally, so values
    agree to rounding (1e-9 relative)"""
    cfg = M.InferenceConfig(1, 2, use_mh_instead_of_pg=True, rejuv_frequency=10 ** 9)
    bad, checked = _param_parity("rents", cfg, ["Obs"], seed=6, max_rows=6000)
    assert checked["slots"] >= 50 and not bad, (checked, bad[:4])


def test_latent_row_move_parity_synthetic_20k():
    """latent-class moves on a 20,000-row synthetic table (hundreds to thousands of referring Records
    per latent row, option lists of thousands of strings): the referrers are summed per DISTINCT
    observed string (score x multiplicity) and long option lists are pruned with the integer bound —
    the installed row and the log-ML still equal the oracle's, which sums referrer by referrer"""
    from tests import test_engine_parity as T
    cfg = M.InferenceConfig(1, 20)
    n = 20000
    model, query, ir, dirty, truth, o, e = _setup_synth(cfg, n, H=1024)
    bad = T._latent_parity(model, query, ir, o, e, 11, 1, ["Hospital"], per_class=5)
    bad += T._latent_parity(model, query, ir, o, e, 11, 1, ["Place", "County"], per_class=3)
    bad += T._latent_parity(model, query, ir, o, e, 11, 1, ["Measure", "HospitalType"], per_class=1)
    assert not bad, (len(bad), bad[:3])
    # and with the pruned path switched off the same moves select the same rows (exhaustive grouped sums)
    e.set_option("prune", 0)
    bad = T._latent_parity(model, query, ir, o, e, 11, 1, ["Place"], per_class=2)
    assert not bad, (len(bad), bad[:3])
