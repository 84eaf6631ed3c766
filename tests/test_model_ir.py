"""The Python host mirror must reproduce the reference builder's IR (SURVEY App. A:
vertex numbering, blocks, plans, query maps — derived from src/dsl/builder.jl)."""
import numpy as np

from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment


def kinds(cm):
    out = {}
    for n in cm.nodes:
        k = "Sub" if isinstance(n, M.SubmodelNode) else type(n).__name__
        out[k] = out.get(k, 0) + 1
    return out


def test_hospital_ir(hospital):
    model, query, dirty, clean, ir, obs = hospital
    assert model.class_order == ["County", "Place", "Condition", "Measure", "HospitalType", "Hospital", "Record"]
    nv = {c: len(model.classes[c].nodes) for c in model.class_order}
    assert nv == {"County": 11, "Place": 17, "Condition": 5, "Measure": 14, "HospitalType": 5, "Hospital": 53, "Record": 67}
    rec = model.classes["Record"]
    assert [len(b) for b in rec.blocks] == [49, 15]
    assert kinds(rec) == {"ForeignKeyNode": 2, "Sub": 49, "RandomChoiceNode": 15, "JuliaNode": 1}
    assert query.obsmap == {"ProviderNumber": 43, "HospitalName": 44, "HospitalType": 51, "HospitalOwner": 52, "Address1": 45,
                            "PhoneNumber": 50, "EmergencyService": 42, "City": 46, "CountyName": 49, "State": 47, "ZipCode": 48,
                            "Condition": 65, "MeasureCode": 63, "MeasureName": 64, "Stateavg": 67}
    assert query.cleanmap["State"] == 8 and query.cleanmap["Stateavg"] == 66 and query.cleanmap["Condition"] == 62
    # block 2 plan: 53 -> [54 -> 55 -> {63, 66 -> 67}], [56 -> 57 -> 64], [58 -> ... -> 62 -> 65]
    p = rec.plans[1]
    assert [s.idx for s in p.steps] == [53]
    assert [s.idx for s in p.steps[0].rest.steps] == [54, 56, 58]
    county = model.classes["County"]
    assert sorted(len(p) for p in county.incoming_references) == [1, 2, 3]


def test_rents_and_flights_ir():
    model, query, *_ = load_experiment("rents", max_rows=500)
    assert len(model.classes["Obs"].nodes) == 20 and [len(b) for b in model.classes["Obs"].blocks] == [18]
    assert model.classes["County"].hash_keys == [2]
    assert query.obsmap == {"CountyKey": 4, "County": 12, "State": 10, "Room Type": 14, "Monthly Rent": 19}
    assert query.cleanmap["Monthly Rent"] == 20
    model, query, *_ = load_experiment("flights", max_rows=500)
    assert len(model.classes["Obs"].nodes) == 28 and [len(b) for b in model.classes["Obs"].blocks] == [13, 6, 8]
    assert len(model.classes["Flight"].nodes) == 21 and [len(b) for b in model.classes["Flight"].blocks] == [9, 12]
    assert model.classes["Flight"].hash_keys == [4]
    assert query.obsmap["sched_dep_time"] == 22 and query.cleanmap["sched_dep_time"] == 8


def test_flat_ir_roundtrip(hospital):
    model, query, dirty, clean, ir, obs = hospital
    a = ir._arrays
    assert a["class_voff"][-1] == 172 and ir.n_blocks == 8
    # explicit-missing vs absent: hospital has no missing queried cells
    assert obs.n_rows == 1000 and obs.n_cols == 15
    # every string of the data is in the dictionary
    assert ir.string_id["birmingham"] >= 0
