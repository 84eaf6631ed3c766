"""Flights program — reference `experiments/flights/run.jl:5-45`, `load_data.jl:1-14`."""
from __future__ import annotations

from typing import Dict, List

from ..model import (MaybeSwap, PARAM_PROB, PCleanModelBuilder, StringPrior, TimePrior, make_query)
from .data import unique_in_order
from .hospital import const

TIME_FIELDS = ["sched_dep_time", "sched_arr_time", "act_dep_time", "act_arr_time"]

FLIGHTS_QUERY = [  # run.jl:38-45
    ("sched_dep_time", "flight.sdt", "sdt"),
    ("sched_arr_time", "flight.sat", "sat"),
    ("act_dep_time", "flight.adt", "adt"),
    ("act_arr_time", "flight.aat", "aat"),
    ("flight", "flight.flight_id"),
    ("src", "src.name"),
]


def build_flights(dirty: Dict[str, List]):
    flight_ids = unique_in_order(dirty["flight"])
    websites = unique_in_order(dirty["src"])
    times: Dict[str, List[str]] = {f"{fl}-{f}": [] for fl in flight_ids for f in TIME_FIELDS}
    n = len(dirty["flight"])
    for i in range(n):
        for f in TIME_FIELDS:
            v = dirty[f][i]
            if v is not None:
                lst = times[f"{dirty['flight'][i]}-{f}"]
                if v not in lst:
                    lst.append(v)

    b = PCleanModelBuilder()
    b.add_new_class("TrackingWebsite")
    b.add_choice_node("TrackingWebsite", "name", StringPrior, [const(2), const(30), const(websites)])
    b.finish_class("TrackingWebsite")

    b.add_new_class("Flight")
    b.begin_block("Flight")
    b.add_choice_node("Flight", "flight_id", StringPrior, [const(10), const(20), const(flight_ids)])
    b.add_guaranteed("Flight", "flight_id")
    b.end_block()
    for name, fld in zip(["sdt", "sat", "adt", "aat"], TIME_FIELDS):
        b.add_choice_node("Flight", name, TimePrior,
                          [(["flight_id"], lambda fid, fld=fld: times[f"{fid}-{fld}"])])
    b.finish_class("Flight")

    b.add_new_class("Obs")
    b.add_indexed_parameter("Obs", "error_probs", PARAM_PROB, 10.0, 50.0)
    b.begin_block("Obs")
    b.add_foreign_key("Obs", "flight", "Flight")
    b.end_block()
    b.add_foreign_key("Obs", "src", "TrackingWebsite")
    # error_prob = lowercase(src.name) == lowercase(flight.flight_id[1:2]) ? 1e-5 : error_probs[src.name]
    b.add_julia_node("Obs", "error_prob", ["src.name", "flight.flight_id", "error_probs"],
                     lambda name, fid, probs: 1e-5 if name.lower() == fid[:2].lower() else probs[name])
    b.begin_block("Obs")
    for name, fld in zip(["sdt", "sat", "adt", "aat"], TIME_FIELDS):
        b.add_choice_node("Obs", name, MaybeSwap,
                          [f"flight.{name}",
                           (["flight.flight_id"], lambda fid, fld=fld: times[f"{fid}-{fld}"]),
                           "error_prob"])
    b.end_block()
    b.finish_class("Obs")

    model = b.finish_model()
    query = make_query(model, "Obs", FLIGHTS_QUERY)
    return model, query
