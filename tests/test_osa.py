"""The bit-parallel OSA distance (Hyyrö 2003; single- and multi-word) equals the plain DP the
oracle restates from StringDistances' DamerauLevenshtein — on random strings, near-duplicates,
empty strings and strings longer than 64 / 128 / 192 symbols, with 3-segment texts (joins)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def osa():
    so = os.path.join(HERE, "_osa_host.so")
    src = os.path.join(HERE, "osa_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.osa_bitpar.argtypes = [C.c_char_p, C.c_int] + [C.c_char_p, C.c_int] * 3
    return L


def dp(a, b):
    n, m = len(a), len(b)
    if not n:
        return m
    if not m:
        return n
    pp, p = None, list(range(m + 1))
    for i in range(1, n + 1):
        c = [i] + [0] * m
        for j in range(1, m + 1):
            v = min(p[j] + 1, c[j - 1] + 1, p[j - 1] + (a[i - 1] != b[j - 1]))
            if i > 1 and j > 1 and a[i - 1] == b[j - 2] and a[i - 2] == b[j - 1]:
                v = min(v, pp[j - 2] + 1)
            c[j] = v
        pp, p = p, c
    return p[m]


def test_bitparallel_matches_dp(osa):
    rng = np.random.default_rng(11)
    for it in range(3000):
        alpha = int(rng.integers(2, 7))
        maxl = [250, 70, 20, 130][it % 4]
        a = bytes(rng.integers(0, alpha, size=int(rng.integers(0, maxl + 1))).astype(np.uint8))
        if it % 2 == 0 and a:
            b = bytearray(a)
            for _ in range(int(rng.integers(0, 6))):
                t, pos = int(rng.integers(0, 4)), int(rng.integers(0, max(1, len(b))))
                if t == 0:
                    b.insert(pos, int(rng.integers(0, alpha)))
                elif t == 1 and b:
                    del b[pos]
                elif t == 2 and pos + 1 < len(b):
                    b[pos], b[pos + 1] = b[pos + 1], b[pos]
                elif b:
                    b[pos] = int(rng.integers(0, alpha))
            b = bytes(b[:250])
        else:
            b = bytes(rng.integers(0, alpha, size=int(rng.integers(0, maxl + 1))).astype(np.uint8))
        c1 = int(rng.integers(0, len(b) + 1)); c2 = int(rng.integers(c1, len(b) + 1))
        got = osa.osa_bitpar(a, len(a), b[:c1], c1, b[c1:c2], c2 - c1, b[c2:], len(b) - c2)
        assert got == dp(a, b), (a, b)


def test_matches_oracle_on_dataset_strings(osa, hospital):
    from oracle import Oracle
    from pclean_b200.host_fixture import model as M
    model, query, dirty, clean, ir, obs = hospital
    o = Oracle(ir, M.InferenceConfig(1, 2), seed=0)
    rng = np.random.default_rng(3)
    strs = [s for s in ir.strings if s]
    for _ in range(1500):
        a, b = strs[int(rng.integers(0, len(strs)))], strs[int(rng.integers(0, len(strs)))]
        ea, eb = a.encode("latin-1", "replace"), b.encode("latin-1", "replace")
        assert osa.osa_bitpar(ea, len(ea), eb, len(eb), b"", 0, b"", 0) == o.edit_distance(a, b)
