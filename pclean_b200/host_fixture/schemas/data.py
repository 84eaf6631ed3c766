"""CSV loading with the typing the reference scripts rely on (CSV.jl inference):
empty cell → missing (None); a column is numeric only if asked for."""
from __future__ import annotations

import csv
from typing import Dict, Iterable, List, Optional


def load_csv(path: str, float_cols: Iterable[str] = ()) -> Dict[str, List]:
    float_cols = set(float_cols)
    with open(path, newline="", encoding="utf-8") as fh:
        rd = csv.reader(fh)
        header = next(rd)
        cols: Dict[str, List] = {h: [] for h in header}
        for row in rd:
            for h, cell in zip(header, row):
                if cell == "":
                    cols[h].append(None)
                elif h in float_cols:
                    cols[h].append(float(cell))
                else:
                    cols[h].append(cell)
    return cols


def unique_in_order(values: Iterable) -> List:
    """Julia `unique` keeps first-occurrence order; `remove_missing` drops missing."""
    seen = set()
    out = []
    for v in values:
        if v is None or v in seen:
            continue
        seen.add(v)
        out.append(v)
    return out


def n_rows(data: Dict[str, List]) -> int:
    return len(next(iter(data.values())))
