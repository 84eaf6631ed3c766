"""Rents program — reference `experiments/rents/run.jl:5-35`, `load_data.jl:1-18`."""
from __future__ import annotations

from typing import Dict, List

from ..model import (AddTypos, ChooseProportionally, ChooseUniformly, PARAM_MEAN, PARAM_PROPORTIONS,
                     PCleanModelBuilder, StringPrior, TransformedGaussian, Transformation, Unmodeled,
                     make_query)
from .data import unique_in_order
from .hospital import const

ROOM_TYPES = ["studio", "1br", "2br", "3br", "4br"]                     # load_data.jl:18
UNITS = [Transformation(1.0), Transformation(1000.0)]                   # run.jl:5-6
BUILTIN_ROUND_BACKWARD = ("round_backward",)   # corrected = round(unit.backward(rent)), run.jl:25

RENTS_QUERY = [  # run.jl:29-35
    ("CountyKey", "county.countykey"),
    ("County", "county.name", "county_name"),
    ("State", "county.state"),
    ("Room Type", "br"),
    ("Monthly Rent", "corrected", "rent"),
]


def county_key(name: str) -> str:
    """load_data.jl:9 — first character + last character of the first word."""
    return f"{name[0]}{name.split()[0][-1]}"


def add_county_key(dirty: Dict[str, List]) -> None:
    dirty["CountyKey"] = [county_key(x) for x in dirty["County"]]


def build_rents(dirty: Dict[str, List]):
    if "CountyKey" not in dirty:
        add_county_key(dirty)
    poss: Dict[str, List[str]] = {}
    for key, name in zip(dirty["CountyKey"], dirty["County"]):
        lst = poss.setdefault(key, [])
        if name not in lst:
            lst.append(name)
    states = unique_in_order(dirty["State"])

    b = PCleanModelBuilder()
    b.add_new_class("County")
    b.add_basic_parameter("County", "state_pops", PARAM_PROPORTIONS)
    b.add_choice_node("County", "countykey", Unmodeled, [])
    b.add_guaranteed("County", "countykey")
    b.add_choice_node("County", "name", StringPrior,
                      [const(10), const(35), (["countykey"], lambda k: poss[k])])
    b.add_choice_node("County", "state", ChooseProportionally, [const(states), "state_pops"])
    b.finish_class("County")

    b.add_new_class("Obs")
    b.add_indexed_parameter("Obs", "avg_rent", PARAM_MEAN, 1500, 1000)
    b.add_foreign_key("Obs", "county", "County")
    b.add_choice_node("Obs", "county_name", AddTypos, ["county.name", const(2)])
    b.add_choice_node("Obs", "br", ChooseUniformly, [const(ROOM_TYPES)])
    b.add_choice_node("Obs", "unit", ChooseUniformly, [const(UNITS)])
    b.add_julia_node("Obs", "rent_base", ["avg_rent", "county.state", "county.countykey", "br"],
                     lambda avg_rent, state, key, br: avg_rent[f"{state}_{key}_{br}"])
    b.add_choice_node("Obs", "rent", TransformedGaussian, ["rent_base", const(150.0), "unit"])
    b.add_julia_node("Obs", "corrected", ["unit", "rent"],
                     lambda unit, rent: float(round(unit.backward(rent))), BUILTIN_ROUND_BACKWARD)
    b.finish_class("Obs")

    model = b.finish_model()
    query = make_query(model, "Obs", RENTS_QUERY)
    return model, query


# ------------------------------------------------------------------------------------------
# the rents schema widened to five AddTypos string columns (BASELINE.json configs[4], SURVEY 8d
# "R10M": county name + four more name-like attributes of the county, each observed with
# AddTypos(max_typos = 2)); same structure as experiments/rents/run.jl otherwise
# ------------------------------------------------------------------------------------------
EXTRA_ATTRS = ["seat", "region", "office", "clerk"]
EXTRA_COLUMNS = {"seat": "County Seat", "region": "Region", "office": "Office", "clerk": "Clerk"}

RENTS5_QUERY = RENTS_QUERY + [(EXTRA_COLUMNS[a], f"county.{a}", f"county_{a}") for a in EXTRA_ATTRS]


def build_rents5(dirty: Dict[str, List]):
    if "CountyKey" not in dirty:
        add_county_key(dirty)
    cols = {"name": "County", **EXTRA_COLUMNS}
    poss: Dict[str, Dict[str, List[str]]] = {a: {} for a in cols}
    for a, col in cols.items():
        seen = set()
        for key, val in zip(dirty["CountyKey"], dirty[col]):
            if val is None or (key, val) in seen:
                continue
            seen.add((key, val))
            poss[a].setdefault(key, []).append(val)
    states = unique_in_order(dirty["State"])

    b = PCleanModelBuilder()
    b.add_new_class("County")
    b.add_basic_parameter("County", "state_pops", PARAM_PROPORTIONS)
    b.add_choice_node("County", "countykey", Unmodeled, [])
    b.add_guaranteed("County", "countykey")
    b.add_choice_node("County", "name", StringPrior, [const(10), const(35), (["countykey"], lambda k: poss["name"][k])])
    for a in EXTRA_ATTRS:
        b.add_choice_node("County", a, StringPrior, [const(10), const(35), (["countykey"], lambda k, a=a: poss[a][k])])
    b.add_choice_node("County", "state", ChooseProportionally, [const(states), "state_pops"])
    b.finish_class("County")

    b.add_new_class("Obs")
    b.add_indexed_parameter("Obs", "avg_rent", PARAM_MEAN, 1500, 1000)
    b.add_foreign_key("Obs", "county", "County")
    b.add_choice_node("Obs", "county_name", AddTypos, ["county.name", const(2)])
    for a in EXTRA_ATTRS:
        b.add_choice_node("Obs", f"county_{a}", AddTypos, [f"county.{a}", const(2)])
    b.add_choice_node("Obs", "br", ChooseUniformly, [const(ROOM_TYPES)])
    b.add_choice_node("Obs", "unit", ChooseUniformly, [const(UNITS)])
    b.add_julia_node("Obs", "rent_base", ["avg_rent", "county.state", "county.countykey", "br"],
                     lambda avg_rent, state, key, br: avg_rent[f"{state}_{key}_{br}"])
    b.add_choice_node("Obs", "rent", TransformedGaussian, ["rent_base", const(150.0), "unit"])
    b.add_julia_node("Obs", "corrected", ["unit", "rent"],
                     lambda unit, rent: float(round(unit.backward(rent))), BUILTIN_ROUND_BACKWARD)
    b.finish_class("Obs")

    model = b.finish_model()
    query = make_query(model, "Obs", RENTS5_QUERY)
    return model, query
