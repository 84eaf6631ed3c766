"""Shared experiment plumbing: load a shipped dataset, build its program, flatten it.
(Callers of the hot path — `experiments/*/run.jl` — not the path itself.)"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

from . import model as M
from ..lowering import FlatIR
from .schemas import build_flights, build_hospital, build_rents, load_csv
from .schemas.rents import add_county_key

DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "datasets")


def load_experiment(name: str, data_dir: str = DATA_DIR, max_rows: int = None):
    """Returns (model, query, dirty, clean, ir, obs)."""
    if name == "hospital":
        dirty = load_csv(os.path.join(data_dir, "hospital_dirty.csv"))
        clean = load_csv(os.path.join(data_dir, "hospital_clean.csv"))
        builder = build_hospital
    elif name == "rents":
        dirty = load_csv(os.path.join(data_dir, "rents_dirty.csv"), float_cols=["Monthly Rent"])
        clean = load_csv(os.path.join(data_dir, "rents_clean.csv"), float_cols=["Monthly Rent"])
        add_county_key(dirty)
        builder = build_rents
    elif name == "flights":
        dirty = load_csv(os.path.join(data_dir, "flights_dirty.csv"))
        clean = load_csv(os.path.join(data_dir, "flights_clean.csv"))
        builder = build_flights
    else:
        raise ValueError(name)
    if max_rows is not None:
        dirty = {k: v[:max_rows] for k, v in dirty.items()}
        clean = {k: v[:max_rows] for k, v in clean.items()}
    model, query = builder(dirty)
    ds = M.ObservedDataset(query, dirty)
    ir = FlatIR(model, [ds])
    obs = ir.encode_observations(ds)
    return model, query, dirty, clean, ir, obs


def cleaned_columns(query: M.Query, decode_column) -> Dict[str, List]:
    """ours[col] = [trace.tables[cls].rows[i][cleanmap[col]] ...] via a column decoder."""
    return {col: decode_column(v - 1) for col, v in query.cleanmap.items()}
