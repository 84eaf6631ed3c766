"""Hospital program — reference `experiments/hospital/run.jl:5-74`, `load_data.jl:1-19`."""
from __future__ import annotations

from typing import Dict, List

from ..model import (AddTypos, ChooseProportionally, ChooseUniformly, PARAM_PROPORTIONS,
                     PCleanModelBuilder, StringPrior, make_query)
from .data import unique_in_order

HOSPITAL_QUERY = [  # run.jl:58-74  (column, clean expr, dirty expr)
    ("ProviderNumber", "hosp.provider", "provider"),
    ("HospitalName", "hosp.name", "name"),
    ("HospitalType", "hosp.type.desc", "type"),
    ("HospitalOwner", "hosp.owner", "owner"),
    ("Address1", "hosp.addr", "addr"),
    ("PhoneNumber", "hosp.phone", "phone"),
    ("EmergencyService", "hosp.service", "service"),
    ("City", "hosp.loc.city", "city"),
    ("CountyName", "hosp.loc.county.county", "county"),
    ("State", "hosp.loc.county.state", "state"),
    ("ZipCode", "hosp.zip", "zip"),
    ("Condition", "metric.condition.desc", "condition"),
    ("MeasureCode", "metric.code", "code"),
    ("MeasureName", "metric.name", "mname"),
    ("Stateavg", "stateavg", "stateavg_obs"),
]


def const(value):
    """A non-name argument expression becomes a zero-argument JuliaNode (syntax.jl:132-135)."""
    return ([], lambda value=value: value)


def build_hospital(dirty: Dict[str, List]):
    """Returns (model, query).  `possibilities[col]` = unique non-missing dirty values
    (load_data.jl:18-19)."""
    poss = {c: unique_in_order(v) for c, v in dirty.items()}
    b = PCleanModelBuilder()

    b.add_new_class("County")
    b.add_basic_parameter("County", "state_proportions", PARAM_PROPORTIONS)
    b.add_choice_node("County", "state", ChooseProportionally, [const(poss["State"]), "state_proportions"])
    b.add_choice_node("County", "county", StringPrior, [const(3), const(30), const(poss["CountyName"])])
    b.finish_class("County")

    b.add_new_class("Place")
    b.add_foreign_key("Place", "county", "County")
    b.add_choice_node("Place", "city", StringPrior, [const(3), const(30), const(poss["City"])])
    b.finish_class("Place")

    b.add_new_class("Condition")
    b.add_choice_node("Condition", "desc", StringPrior, [const(5), const(35), const(poss["Condition"])])
    b.finish_class("Condition")

    b.add_new_class("Measure")
    b.add_choice_node("Measure", "code", ChooseUniformly, [const(poss["MeasureCode"])])
    b.add_choice_node("Measure", "name", ChooseUniformly, [const(poss["MeasureName"])])
    b.add_foreign_key("Measure", "condition", "Condition")
    b.finish_class("Measure")

    b.add_new_class("HospitalType")
    b.add_choice_node("HospitalType", "desc", StringPrior, [const(10), const(30), const(poss["HospitalType"])])
    b.finish_class("HospitalType")

    b.add_new_class("Hospital")
    b.add_basic_parameter("Hospital", "owner_dist", PARAM_PROPORTIONS)
    b.add_basic_parameter("Hospital", "service_dist", PARAM_PROPORTIONS)
    b.add_foreign_key("Hospital", "loc", "Place")
    b.add_foreign_key("Hospital", "type", "HospitalType")
    b.add_choice_node("Hospital", "provider", ChooseUniformly, [const(poss["ProviderNumber"])])
    b.add_choice_node("Hospital", "name", StringPrior, [const(3), const(50), const(poss["HospitalName"])])
    b.add_choice_node("Hospital", "addr", StringPrior, [const(10), const(30), const(poss["Address1"])])
    b.add_choice_node("Hospital", "phone", StringPrior, [const(10), const(10), const(poss["PhoneNumber"])])
    b.add_choice_node("Hospital", "owner", ChooseProportionally, [const(poss["HospitalOwner"]), "owner_dist"])
    b.add_choice_node("Hospital", "zip", ChooseUniformly, [const(poss["ZipCode"])])
    b.add_choice_node("Hospital", "service", ChooseProportionally, [const(poss["EmergencyService"]), "service_dist"])
    b.finish_class("Hospital")

    b.add_new_class("Record")
    b.begin_block("Record")
    b.add_foreign_key("Record", "hosp", "Hospital")
    b.add_choice_node("Record", "service", AddTypos, ["hosp.service"])
    b.add_choice_node("Record", "provider", AddTypos, ["hosp.provider"])
    b.add_choice_node("Record", "name", AddTypos, ["hosp.name"])
    b.add_choice_node("Record", "addr", AddTypos, ["hosp.addr"])
    b.add_choice_node("Record", "city", AddTypos, ["hosp.loc.city"])
    b.add_choice_node("Record", "state", AddTypos, ["hosp.loc.county.state"])
    b.add_choice_node("Record", "zip", AddTypos, ["hosp.zip"])
    b.add_choice_node("Record", "county", AddTypos, ["hosp.loc.county.county"])
    b.add_choice_node("Record", "phone", AddTypos, ["hosp.phone"])
    b.add_choice_node("Record", "type", AddTypos, ["hosp.type.desc"])
    b.add_choice_node("Record", "owner", AddTypos, ["hosp.owner"])
    b.end_block()
    b.begin_block("Record")
    b.add_foreign_key("Record", "metric", "Measure")
    b.add_choice_node("Record", "code", AddTypos, ["metric.code"])
    b.add_choice_node("Record", "mname", AddTypos, ["metric.name"])
    b.add_choice_node("Record", "condition", AddTypos, ["metric.condition.desc"])
    # stateavg = "$(hosp.loc.county.state)_$(metric.code)"   (run.jl:52)
    b.add_julia_node("Record", "stateavg", ["hosp.loc.county.state", "metric.code"],
                     lambda state, code: f"{state}_{code}", ("join", "_"))
    b.add_choice_node("Record", "stateavg_obs", AddTypos, ["stateavg"])
    b.end_block()
    b.finish_class("Record")

    model = b.finish_model()
    query = make_query(model, "Record", HOSPITAL_QUERY)
    return model, query
