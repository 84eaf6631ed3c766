"""evaluate_accuracy — the F1 definition of the reference (`src/analysis.jl:36-88`), which is
part of the parity contract.  Host-side, runs once."""
from __future__ import annotations

from typing import Dict, List


def _eq(a, b) -> bool:
    if a is None or b is None:
        return False
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return float(a) == float(b)
    if isinstance(a, str) and isinstance(b, (int, float)):
        return False
    if isinstance(b, str) and isinstance(a, (int, float)):
        return False
    return a == b


def evaluate_accuracy(dirty: Dict[str, List], clean: Dict[str, List], ours: Dict[str, List],
                      cleanmap_cols: List[str]) -> dict:
    """`ours[col][i]` is the cleaned value of `cleanmap[col]` for row i (only for queried
    columns).  Mirrors analysis.jl:36-88 line for line."""
    total_errors = total_changed = total_cleaned = total_imputed = total_imputed_correctly = 0
    n = len(next(iter(dirty.values())))
    for i in range(n):
        for col in clean.keys():
            if col not in dirty:
                continue
            d, c = dirty[col][i], clean[col][i]
            if d is None:
                if col in cleanmap_cols and c is not None:
                    total_imputed += 1
                    if _eq(ours[col][i], c):
                        total_imputed_correctly += 1
                continue
            if not _eq(d, c):
                total_errors += 1
            if col in cleanmap_cols:
                o = ours[col][i]
                if not _eq(o, d):
                    total_changed += 1
                    if _eq(o, c):
                        total_cleaned += 1
    denom_p = total_changed + total_imputed
    denom_r = total_errors + total_imputed
    precision = (total_cleaned + total_imputed_correctly) / denom_p if denom_p else float("nan")
    recall = (total_cleaned + total_imputed_correctly) / denom_r if denom_r else float("nan")
    f1 = 2.0 / (1 / precision + 1 / recall) if precision and recall and precision == precision else float("nan")
    return dict(f1=f1, errors=total_errors, changed=total_changed, cleaned=total_cleaned, precision=precision,
                recall=recall, imputed=total_imputed, correctly_imputed=total_imputed_correctly)
