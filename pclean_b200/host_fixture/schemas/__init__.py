"""The three benchmark programs of the reference (`experiments/{hospital,rents,flights}/run.jl`)
restated as the builder commands `PClean.@model` expands to (`src/dsl/syntax.jl:106-161`).
They are *callers* of the hot path (fixtures), not part of it."""
from .hospital import build_hospital            # noqa: F401
from .rents import build_rents                  # noqa: F401
from .flights import build_flights              # noqa: F401
from .data import load_csv, unique_in_order     # noqa: F401
