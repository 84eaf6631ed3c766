"""Host-side FIXTURES, not product: stand-ins for the parts of the reference that stay Julia and are
out of scope — the builder commands `PClean.@model` expands to (`model.py` ↔
`src/dsl/builder.jl`), the three shipped programs (`schemas/` ↔ `experiments/*/run.jl`), their data
loading (`experiments.py`), `evaluate_accuracy` (`analysis.py` ↔ `src/analysis.jl:36-88`) and the
synthetic workload generators (`synth.py`).  They exist because the build image has no Julia and the
engine needs a model IR, datasets and an accuracy read-out to be tested and benchmarked; a Julia host
uses `julia/PCleanB200.jl` instead."""
