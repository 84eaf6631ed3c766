"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE — see pclean_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg
may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from pclean_b200.lowering import Config, FlatIR, ModelIR, Observations, VALUE_DTYPE, Value

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpclean_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pclean_oracle.cpp")
    deps = [src, os.path.join(_HERE, "..", "include", "pclean_b200.h"), os.path.join(_HERE, "..", "include", "pclean_rng.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(_LIB_PATH) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _LIB_PATH, src], cwd=_HERE)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(ModelIR), C.POINTER(Config), C.c_uint64]
        L.oracle_clone.restype = C.c_void_p
        L.oracle_clone.argtypes = [C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error.argtypes = [C.c_void_p]
        L.oracle_set_config.argtypes = [C.c_void_p, C.POINTER(Config)]
        L.oracle_set_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_set_true_damerau.argtypes = [C.c_void_p, C.c_int]
        L.oracle_load_observations.argtypes = [C.c_void_p, C.POINTER(Observations)]
        for name in ("oracle_initialize_trace", "oracle_sweep", "oracle_run_inference", "oracle_begin_sweep"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_sweep_class.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.oracle_initialize_prefix.argtypes = [C.c_void_p, C.c_int64]
        L.oracle_row_move.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_install_table.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.c_void_p]
        L.oracle_install_obs_rows.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.c_int64]
        L.oracle_bump_refcount.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.oracle_reset_tables.argtypes = [C.c_void_p]
        L.oracle_table_size.restype = C.c_int64
        L.oracle_table_size.argtypes = [C.c_void_p, C.c_int]
        L.oracle_table_keys.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_get_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_void_p]
        L.oracle_get_py.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.oracle_set_py.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.oracle_string_count.argtypes = [C.c_void_p]
        L.oracle_get_string.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        L.oracle_intern_string.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
        L.oracle_param_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.oracle_param_set.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.oracle_n_slots.argtypes = [C.c_void_p]
        L.oracle_edit_distance.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_addtypos.restype = C.c_double
        L.oracle_addtypos.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_stringprior.restype = C.c_double
        L.oracle_stringprior.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_logdensity.restype = C.c_double
        L.oracle_logdensity.argtypes = [C.c_void_p, C.c_int, C.POINTER(Value), C.c_int, C.POINTER(Value)]
        L.oracle_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_crp_logprior.restype = C.c_double
        L.oracle_crp_logprior.argtypes = [C.c_int64, C.c_double, C.c_double, C.c_int64]
        L.oracle_logsumexp.restype = C.c_double
        L.oracle_logsumexp.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.oracle_set_epochs.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_resample_class.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


class Oracle:
    """The reference path restated on the CPU: initialize_trace / run_inference! and helpers."""

    def __init__(self, ir: FlatIR, config, seed: int = 0, _handle=None):
        self.ir = ir
        self.L = lib()
        self.config = config
        if _handle is not None:
            self.h = _handle
            return
        cfg = Config.from_config(config)
        self._cir = ir.as_ctypes()
        self.h = self.L.oracle_create(C.byref(self._cir), C.byref(cfg), C.c_uint64(seed))
        if not self.h:
            raise OracleError("oracle_create failed")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(self.L.oracle_last_error(self.h).decode())

    def clone(self) -> "Oracle":
        return Oracle(self.ir, self.config, _handle=self.L.oracle_clone(self.h))

    def set_config(self, config):
        self.config = config
        cfg = Config.from_config(config)
        self.L.oracle_set_config(self.h, C.byref(cfg))

    def set_seed(self, seed: int):
        self.L.oracle_set_seed(self.h, C.c_uint64(seed))

    def load_observations(self, obs: Observations):
        self._check(self.L.oracle_load_observations(self.h, C.byref(obs)))

    def initialize_trace(self, n_rows: int = None):
        if n_rows is None:
            self._check(self.L.oracle_initialize_trace(self.h))
        else:
            self._check(self.L.oracle_initialize_prefix(self.h, n_rows))

    def sweep(self):
        self._check(self.L.oracle_sweep(self.h))

    def begin_sweep(self):
        self.L.oracle_begin_sweep(self.h)

    def sweep_class(self, cls: int, row_begin: int = 0, row_end: int = -1):
        self._check(self.L.oracle_sweep_class(self.h, cls, row_begin, row_end))

    def run_inference(self):
        self._check(self.L.oracle_run_inference(self.h))

    def row_move(self, cls: int, key: int, n_blocks: int):
        K = self.config.num_particles
        keys = (C.c_int64 * (K * n_blocks))()
        w = (C.c_double * K)()
        sel = C.c_int()
        ml = C.c_double()
        self._check(self.L.oracle_row_move(self.h, cls, key, keys, w, C.byref(sel), C.byref(ml)))
        return (np.array(keys, dtype=np.int64).reshape(K, n_blocks), np.array(w, dtype=np.float64), sel.value, ml.value)

    def install_snapshot(self, ir: FlatIR, model, obs_cls_name: str, snap: dict, n_obs_rows: int = None, bump_to_full: bool = False):
        """Install a trace (same dict `load_trace_from_snapshot` takes).  With `n_obs_rows` only
        a prefix of the observation rows is installed; `bump_to_full` then raises the
        reference counts of the latent rows to what the full assignment implies, so the prefix
        is scored against the full tables (SURVEY §8d CPU-baseline sampling)."""
        self.L.oracle_reset_tables(self.h)
        for name in model.class_order:
            if name == obs_cls_name:
                continue
            keys, cells, s, d = snap["tables"][name]
            keys = np.ascontiguousarray(keys, dtype=np.int64)
            cells = np.ascontiguousarray(cells, dtype=VALUE_DTYPE)
            self._check(self.L.oracle_install_table(self.h, ir.class_index[name], len(keys), cells.shape[0],
                                                    keys.ctypes.data_as(C.POINTER(C.c_int64)), cells.ctypes.data))
            self.set_py(ir.class_index[name], s, d)
        for slot, vals in snap.get("params", {}).items():
            self.param_set(slot, list(vals))
        fks = sorted(snap["assignment"].keys())
        keys = np.ascontiguousarray(np.stack([snap["assignment"][f] for f in fks]), dtype=np.int64)
        n_total = keys.shape[1]
        n = n_total if n_obs_rows is None else min(n_obs_rows, n_total)
        cls = ir.class_index[obs_cls_name]
        self._check(self.L.oracle_install_obs_rows(self.h, cls, n, len(fks), keys.ctypes.data_as(C.POINTER(C.c_int64)), n_total))
        if bump_to_full and n < n_total:
            cm = model.classes[obs_cls_name]
            for fi, f in enumerate(fks):
                target = ir.class_index[cm.nodes[f].target_class]
                uniq, cnt = np.unique(keys[fi, n:], return_counts=True)
                for k, c in zip(uniq, cnt):
                    self._check(self.L.oracle_bump_refcount(self.h, target, int(k), int(c)))

    def table_size(self, cls: int) -> int:
        return self.L.oracle_table_size(self.h, cls)

    def table_keys(self, cls: int):
        n = self.table_size(cls)
        keys = np.zeros(n, dtype=np.int64)
        ref = np.zeros(n, dtype=np.int64)
        self.L.oracle_table_keys(self.h, cls, keys.ctypes.data_as(C.POINTER(C.c_int64)), ref.ctypes.data_as(C.POINTER(C.c_int64)))
        return keys, ref

    def get_cells(self, cls: int, vertices: Sequence[int]) -> np.ndarray:
        """cells[vi, r] for rows in ascending key order (structured VALUE_DTYPE)."""
        n = self.table_size(cls)
        v = np.ascontiguousarray(np.asarray(vertices, dtype=np.int32))
        out = np.zeros((len(v), n), dtype=VALUE_DTYPE)
        self.L.oracle_get_cells(self.h, cls, len(v), v.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data)
        return out

    def get_py(self, cls: int):
        s, d, t = C.c_double(), C.c_double(), C.c_int64()
        self.L.oracle_get_py(self.h, cls, C.byref(s), C.byref(d), C.byref(t))
        return s.value, d.value, t.value

    def set_py(self, cls: int, strength: float, discount: float):
        self.L.oracle_set_py(self.h, cls, strength, discount)

    def string(self, sid: int) -> str:
        n = self.L.oracle_get_string(self.h, sid, 0, None)
        buf = (C.c_uint32 * max(1, n))()
        self.L.oracle_get_string(self.h, sid, n, buf)
        return "".join(chr(c) for c in buf[:n])

    def string_count(self) -> int:
        return self.L.oracle_string_count(self.h)

    def intern(self, s: str) -> int:
        arr = (C.c_uint32 * max(1, len(s)))(*[ord(c) for c in s])
        return self.L.oracle_intern_string(self.h, len(s), arr)

    def param_get(self, slot: int, cap: int = 4096):
        vals = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = self.L.oracle_param_get(self.h, slot, cap, vals, cnt)
        n = min(n, cap)
        return np.array(vals[:n]), np.array(cnt[:n])

    def param_set(self, slot: int, values):
        arr = (C.c_double * len(values))(*values)
        self.L.oracle_param_set(self.h, slot, len(values), arr)

    def set_epochs(self, epoch: int):
        """put every parameter / Pitman-Yor resampling counter at `epoch` (keyed RNG streams)"""
        self.L.oracle_set_epochs(self.h, epoch)

    def resample_class(self, cls: int):
        """the rejuvenation step of pgibbs_sweep! for one class (inference.jl:72-77)"""
        self._check(self.L.oracle_resample_class(self.h, cls))

    def n_slots(self) -> int:
        return self.L.oracle_n_slots(self.h)

    def edit_distance(self, a: str, b: str) -> int:
        return self.L.oracle_edit_distance(self.h, self.intern(a), self.intern(b))

    def addtypos(self, observed: str, word: str, max_typos: int = -1) -> float:
        return self.L.oracle_addtypos(self.h, self.intern(observed), self.intern(word), max_typos)

    def stringprior(self, s: str, minl: int, maxl: int) -> float:
        return self.L.oracle_stringprior(self.h, self.intern(s), minl, maxl)

    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.L.oracle_counters(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"dp_cells": a.value, "typo_evals": b.value, "typo_misses": c.value}

    def decode(self, cell) -> object:
        """structured cell -> python value"""
        from pclean_b200 import lowering as LW
        tag = int(cell["tag"])
        if tag == LW.VAL_STR:
            return self.string(int(cell["i"]))
        if tag == LW.VAL_REAL:
            return float(cell["d"])
        if tag == LW.VAL_INT:
            return int(cell["i"])
        if tag == LW.VAL_KEY:
            return int(cell["d"])
        if tag in (LW.VAL_MISSING, LW.VAL_ABSENT):
            return None
        return (tag, int(cell["i"]), float(cell["d"]))


def export_snapshot(o: "Oracle", ir: FlatIR, model, obs_cls_name: str) -> dict:
    """Dump the oracle's trace in the form `pclean_load_table` / `pclean_load_assignment`
    take (string ids re-interned into `ir`, whose dictionary the engine uploads)."""
    from pclean_b200 import lowering as LW
    from pclean_b200.host_fixture import model as M
    remap = {}

    def fix_strings(cells):
        tags = cells["tag"]
        ids = cells["i"]
        for sid in np.unique(ids[tags == LW.VAL_STR]):
            sid = int(sid)
            if sid not in remap:
                remap[sid] = ir.intern_string(o.string(sid))
        if remap:
            mask = tags == LW.VAL_STR
            ids[mask] = np.vectorize(lambda x: remap[int(x)], otypes=[np.int32])(ids[mask])
        return cells

    snap = {"tables": {}, "assignment": {}, "params": {}, "rowcells": {}}
    for name in model.class_order:
        cls = ir.class_index[name]
        cm = model.classes[name]
        n_normal = sum(1 for n in cm.nodes if not isinstance(n, M.ExternalLikelihoodNode))
        if name == obs_cls_name:
            fks = [v for v, n in enumerate(cm.nodes) if isinstance(n, M.ForeignKeyNode)]
            cells = o.get_cells(cls, fks)
            for k, v in enumerate(fks):
                snap["assignment"][v] = cells[k]["d"].astype(np.int64)
            # local discrete choices of the observed rows (rents: br, unit)
            # local cells the engine keeps per row: enumerated choices (rents br, unit) and cells sampled
            # with random() when the dataset lacks them (flights MaybeSwap observations)
            local = [v for v, n in enumerate(cm.nodes) if isinstance(n, M.RandomChoiceNode)
                     and (M.HAS_DISCRETE_PROPOSAL[n.dist] or n.dist == M.MaybeSwap)]
            if local:
                lc = fix_strings(o.get_cells(cls, local))
                snap["rowcells"] = {v: lc[k].copy() for k, v in enumerate(local)}
            continue
        keys, _ = o.table_keys(cls)
        cells = fix_strings(o.get_cells(cls, list(range(n_normal))))
        s, d, _ = o.get_py(cls)
        snap["tables"][name] = (keys, cells, s, d)
    for slot in range(o.n_slots()):
        vals, _ = o.param_get(slot)
        if len(vals):
            snap["params"][slot] = vals
    ir.refresh()
    return snap
