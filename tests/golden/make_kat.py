"""Recomputes tests/golden/kat.json from the closed forms (no reference execution is possible:
no Julia toolchain).  AddTypos: logpdf(NegativeBinomial(ceil(len/5), 0.9), k) - k log(len) - k log(26)/2."""
import json
import math
import os


def addtypos(k, n):
    r = math.ceil(n / 5)
    return math.lgamma(k + r) - math.lgamma(k + 1) - math.lgamma(r) + r * math.log(0.9) + k * math.log(0.1) - k * math.log(n) - k * math.log(26) / 2


if __name__ == "__main__":
    print(json.dumps({"birmingham k=0": addtypos(0, 10), "k=1": addtypos(1, 10), "al/ak": addtypos(1, 2)}, indent=1))
