"""Python host over the C-ABI engine (plays the Julia shim of INTEGRATION.md).

Mirrors the reference surface: `InferenceConfig`, `initialize_trace`, `run_inference!`
(`src/inference/inference.jl:3,83`) — with the trace living in HBM behind an opaque handle.
The CUDA library is required: importing works anywhere, but creating an `Engine` raises if
`libpclean_b200.so` is missing or no GPU is present (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np

from .host_fixture import model as M
from .lowering import Config, FlatIR, ModelIR, Observations, VALUE_DTYPE, Value, VAL_KEY, VAL_STR

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpclean_b200.so")
_lib = None

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc", "engine.cu")
    deps = [os.path.join(_HERE, "csrc", f) for f in ("engine.cu", "device.cuh", "latent.cuh", "lower.hpp", "osa_bitpar.cuh")]
    deps += [os.path.join(_HERE, "..", "include", f) for f in ("pclean_b200.h", "pclean_rng.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src, "-ldl"]
        subprocess.check_call(cmd, cwd=os.path.join(_HERE, "csrc"))
    return LIB_PATH


class TableSnapshot(C.Structure):
    _fields_ = [("cls", C.c_int32), ("n_rows", C.c_int64), ("n_cols", C.c_int32), ("keys", C.POINTER(C.c_int64)),
                ("cells", C.POINTER(Value)), ("py_strength", C.c_double), ("py_discount", C.c_double)]


class SweepStats(C.Structure):
    _fields_ = [("rows", C.c_int64), ("particles", C.c_int64), ("new_rows", C.c_int64), ("dummy_draws", C.c_int64),
                ("changed_rows", C.c_int64), ("sum_log_ml", C.c_double), ("kernel_ms", C.c_float),
                ("total_ms", C.c_float), ("launches", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "pclean_create", "pclean_destroy", "pclean_last_error", "pclean_version", "pclean_load_model",
    "pclean_load_observations", "pclean_load_table", "pclean_load_assignment", "pclean_set_param_values",
    "pclean_get_param_values", "pclean_init_trace", "pclean_reserve_table", "pclean_sweep", "pclean_run_inference",
    "pclean_row_move_debug", "pclean_download_cells", "pclean_download_assignment", "pclean_download_logweights",
    "pclean_table_size", "pclean_download_table", "pclean_string_count", "pclean_get_string",
    "pclean_addtypos_pairs", "pclean_attach_nccl", "pclean_set_row_shard", "pclean_update_observations",
    "pclean_download_assignment_range", "pclean_download_logweights_range",
    "pclean_load_model_file", "pclean_load_observations_file", "pclean_download_row_flags",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is required; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.pclean_last_error.restype = C.c_char_p
        L.pclean_version.restype = C.c_char_p
        L.pclean_create.argtypes = [C.POINTER(Config), C.c_int32, C.POINTER(C.c_void_p)]
        L.pclean_destroy.argtypes = [C.c_void_p]
        L.pclean_last_error.argtypes = [C.c_void_p]
        L.pclean_load_model.argtypes = [C.c_void_p, C.POINTER(ModelIR)]
        L.pclean_load_observations.argtypes = [C.c_void_p, C.POINTER(Observations)]
        L.pclean_load_model_file.argtypes = [C.c_void_p, C.c_char_p]
        L.pclean_load_observations_file.argtypes = [C.c_void_p, C.c_char_p]
        L.pclean_load_table.argtypes = [C.c_void_p, C.POINTER(TableSnapshot)]
        L.pclean_load_assignment.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.pclean_load_row_cells.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]
        L.pclean_set_param_values.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
        L.pclean_get_param_values.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.pclean_init_trace.argtypes = [C.c_void_p, C.c_uint64]
        L.pclean_reserve_table.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.pclean_sweep.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint32, C.POINTER(SweepStats)]
        L.pclean_run_inference.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(SweepStats)]
        L.pclean_row_move_debug.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.POINTER(C.c_int64),
                                            C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.pclean_download_cells.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int64, C.c_void_p]
        L.pclean_download_assignment.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int64)]
        L.pclean_download_logweights.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)]
        L.pclean_download_assignment_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        L.pclean_download_logweights_range.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
        L.pclean_download_row_flags.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int32)]
        L.pclean_table_size.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
        L.pclean_download_table.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int64)]
        L.pclean_string_count.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.pclean_get_string.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
        L.pclean_addtypos_pairs.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.pclean_debug_distance.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.pclean_attach_nccl.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.pclean_nccl_unique_id.argtypes = [C.c_void_p]
        L.pclean_nccl_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.pclean_set_row_shard.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64]
        L.pclean_latent_move_debug.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.pclean_get_py_params.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pclean_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
        L.pclean_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.pclean_block_metrics.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
        L.pclean_matrix_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.pclean_resync_observations.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.pclean_update_observations.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pclean_b200 error {code}: {msg}")
        self.code = code


class Engine:
    """Opaque PCleanTrace in HBM."""

    def __init__(self, ir, config: M.InferenceConfig, device: int = 0):
        """`ir`: a FlatIR / LoadedIR (anything with as_ctypes()), or the path of a PCLIRv1 file
        (pclean_b200/irfile.py), which the library then reads itself (pclean_load_model_file)"""
        self.L = lib()
        self.ir = ir
        self.config = config
        cfg = Config.from_config(config)
        self.h = C.c_void_p()
        rc = self.L.pclean_create(C.byref(cfg), device, C.byref(self.h))
        if rc != 0:
            raise EngineError(rc, "pclean_create failed (is a CUDA device visible? there is no CPU fallback)")
        if isinstance(ir, (str, bytes, os.PathLike)):
            self._check(self.L.pclean_load_model_file(self.h, os.fsencode(ir)))
        else:
            self._cir = ir.as_ctypes()
            self._check(self.L.pclean_load_model(self.h, C.byref(self._cir)))
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.pclean_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise EngineError(rc, self.L.pclean_last_error(self.h).decode())

    # -- uploads
    def load_observations(self, obs: Observations):
        self._keep.append(obs)
        self._obs = obs
        self._check(self.L.pclean_load_observations(self.h, C.byref(obs)))

    def load_observations_file(self, path):
        self._check(self.L.pclean_load_observations_file(self.h, os.fsencode(path)))

    def load_table(self, cls: int, keys: np.ndarray, cells: np.ndarray, strength: float, discount: float):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        cells = np.ascontiguousarray(cells, dtype=VALUE_DTYPE)
        t = TableSnapshot(cls, len(keys), cells.shape[0], keys.ctypes.data_as(C.POINTER(C.c_int64)),
                          C.cast(cells.ctypes.data, C.POINTER(Value)), strength, discount)
        self._check(self.L.pclean_load_table(self.h, C.byref(t)))

    def load_assignment(self, cls: int, fk_vertices: Sequence[int], keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.int64)      # [n_fk, n_rows]
        v = np.ascontiguousarray(np.asarray(fk_vertices, dtype=np.int32))
        self._check(self.L.pclean_load_assignment(self.h, cls, keys.shape[1], keys.shape[0],
                                                  v.ctypes.data_as(C.POINTER(C.c_int32)), keys.ctypes.data_as(C.POINTER(C.c_int64))))

    def load_row_cells(self, cls: int, vertex: int, values: np.ndarray):
        values = np.ascontiguousarray(values, dtype=VALUE_DTYPE)
        self._check(self.L.pclean_load_row_cells(self.h, cls, vertex, len(values), values.ctypes.data))

    def set_param(self, slot: int, values):
        arr = (C.c_double * len(values))(*values)
        self._check(self.L.pclean_set_param_values(self.h, slot, len(values), arr))

    def get_param(self, slot: int, cap: int = 4096) -> np.ndarray:
        arr = (C.c_double * cap)()
        n = C.c_int32()
        self._check(self.L.pclean_get_param_values(self.h, slot, cap, arr, C.byref(n)))
        return np.array(arr[:min(cap, n.value)])

    # -- hot path
    def reserve_table(self, cls: int, rows: int):
        self._check(self.L.pclean_reserve_table(self.h, cls, rows))

    def init_trace(self, seed: int):
        """initialize_trace (inference.jl:3-58) on the device: batched SMC into empty tables"""
        self._check(self.L.pclean_init_trace(self.h, C.c_uint64(seed)))

    def sweep(self, cls: int, seed: int, sweep_idx: int) -> dict:
        st = SweepStats()
        self._check(self.L.pclean_sweep(self.h, cls, C.c_uint64(seed), sweep_idx, C.byref(st)))
        return st.as_dict()

    def run_inference(self, seed: int) -> dict:
        st = SweepStats()
        self._check(self.L.pclean_run_inference(self.h, C.c_uint64(seed), C.byref(st)))
        return st.as_dict()

    def row_move_debug(self, cls: int, row: int, seed: int, sweep_idx: int, n_blocks: int):
        K = self.config.num_particles
        keys = (C.c_int64 * (K * n_blocks))()
        w = (C.c_double * K)()
        sel = C.c_int32()
        ml = C.c_double()
        self._check(self.L.pclean_row_move_debug(self.h, cls, row, C.c_uint64(seed), sweep_idx, keys, w, C.byref(sel), C.byref(ml)))
        return np.array(keys, dtype=np.int64).reshape(K, n_blocks), np.array(w), sel.value, ml.value

    def latent_move_debug(self, cls: int, key: int, seed: int, sweep_idx: int, n_normal: int):
        out = np.zeros(n_normal, dtype=VALUE_DTYPE)
        sel = C.c_int32()
        ml = C.c_double()
        self._check(self.L.pclean_latent_move_debug(self.h, cls, key, C.c_uint64(seed), sweep_idx, out.ctypes.data, C.byref(sel), C.byref(ml)))
        return out, sel.value, ml.value

    # -- results
    def download_cells(self, cls: int, vertices: Sequence[int], n_rows: int) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(vertices, dtype=np.int32))
        out = np.zeros((len(v), n_rows), dtype=VALUE_DTYPE)
        self._check(self.L.pclean_download_cells(self.h, cls, len(v), v.ctypes.data_as(C.POINTER(C.c_int32)), n_rows, out.ctypes.data))
        return out

    def download_assignment(self, cls: int, fk_vertex: int, n_rows: int) -> np.ndarray:
        out = np.zeros(n_rows, dtype=np.int64)
        self._check(self.L.pclean_download_assignment(self.h, cls, fk_vertex, n_rows, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def download_logweights(self, cls: int, n_rows: int) -> np.ndarray:
        out = np.zeros(n_rows, dtype=np.float64)
        self._check(self.L.pclean_download_logweights(self.h, cls, n_rows, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def download_assignment_range(self, cls: int, fk_vertex: int, begin: int, end: int) -> np.ndarray:
        out = np.zeros(max(0, end - begin), dtype=np.int64)
        self._check(self.L.pclean_download_assignment_range(self.h, cls, fk_vertex, begin, end, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def download_logweights_range(self, cls: int, begin: int, end: int) -> np.ndarray:
        out = np.zeros(max(0, end - begin), dtype=np.float64)
        self._check(self.L.pclean_download_logweights_range(self.h, cls, begin, end, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def download_row_flags(self, cls: int, begin: int, end: int) -> np.ndarray:
        out = np.zeros(max(0, end - begin), dtype=np.int32)
        self._check(self.L.pclean_download_row_flags(self.h, cls, begin, end, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def table_size(self, cls: int) -> int:
        n = C.c_int64()
        self._check(self.L.pclean_table_size(self.h, cls, C.byref(n)))
        return n.value

    def download_table(self, cls: int, n_cols: int = None):
        """keys and reference counts of a latent table; with `n_cols` (the class's non-external vertices) also its cells [n_cols][n]"""
        n = self.table_size(cls)
        keys = np.zeros(n, dtype=np.int64)
        ref = np.zeros(n, dtype=np.int32)
        got = C.c_int64()
        cells = np.zeros((n_cols, n), dtype=VALUE_DTYPE) if n_cols else None
        self._check(self.L.pclean_download_table(self.h, cls, n, keys.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 ref.ctypes.data_as(C.POINTER(C.c_int32)), cells.ctypes.data if n_cols else None, C.byref(got)))
        return (keys, ref, cells) if n_cols else (keys, ref)

    def string(self, sid: int) -> str:
        n = C.c_int32()
        self._check(self.L.pclean_get_string(self.h, sid, 0, None, C.byref(n)))
        buf = (C.c_uint32 * max(1, n.value))()
        self._check(self.L.pclean_get_string(self.h, sid, n.value, buf, C.byref(n)))
        return "".join(chr(c) for c in buf[:n.value])

    def addtypos_pairs(self, observed_ids, clean_ids, max_typos: int = -1):
        a = np.ascontiguousarray(np.asarray(observed_ids, dtype=np.int32))
        b = np.ascontiguousarray(np.asarray(clean_ids, dtype=np.int32))
        d = np.zeros(len(a), dtype=np.int32)
        l = np.zeros(len(a), dtype=np.float64)
        self._check(self.L.pclean_addtypos_pairs(self.h, len(a), a.ctypes.data_as(C.POINTER(C.c_int32)), b.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 max_typos, d.ctypes.data_as(C.POINTER(C.c_int32)), l.ctypes.data_as(C.POINTER(C.c_double))))
        return d, l

    def debug_distance(self, obs_col: int, u: int, table: int, col: int, slot: int) -> int:
        out = C.c_int32()
        self._check(self.L.pclean_debug_distance(self.h, obs_col, u, table, col, slot, C.byref(out)))
        return out.value

    def get_py_params(self, cls: int):
        a, b = C.c_double(), C.c_double()
        self._check(self.L.pclean_get_py_params(self.h, cls, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_counters(self):
        out = (C.c_int32 * 32)()
        self._check(self.L.pclean_debug_counters(self.h, out))
        return list(out)

    def set_option(self, name: str, value: int):
        self._check(self.L.pclean_set_option(self.h, name.encode(), value))

    def block_metrics(self, block: int) -> dict:
        out = (C.c_double * 6)()
        self._check(self.L.pclean_block_metrics(self.h, block, out))
        return {"kernel_ms": out[0], "distance_bytes_per_row": out[1], "elements_per_row": out[2], "terms_per_row": out[3],
                "root_candidates": out[4], "root_terms": out[5]}

    def matrix_bytes(self) -> int:
        n = C.c_int64()
        self._check(self.L.pclean_matrix_bytes(self.h, C.byref(n)))
        return n.value

    def resync_observations(self) -> int:
        n = C.c_int64()
        self._check(self.L.pclean_resync_observations(self.h, C.byref(n)))
        return n.value

    def update_observations(self, sid_cols, real_cols, begin: int, end: int) -> int:
        """Observed cells of rows [begin, end) from the caller's encoded columns: `sid_cols[c]` an int32
        array of dictionary ids (or None), `real_cols[c]` a float64 array (or None), one entry per
        dataset column.  Returns the bytes copied host -> device."""
        nc = len(sid_cols)
        sp = (C.c_void_p * nc)(*[a.ctypes.data if a is not None else None for a in sid_cols])
        rp = (C.c_void_p * nc)(*[a.ctypes.data if a is not None else None for a in real_cols])
        n = C.c_int64()
        self._check(self.L.pclean_update_observations(self.h, nc, sp, rp, begin, end, C.byref(n)))
        return n.value

    def set_row_shard(self, cls: int, begin: int, end: int):
        self._check(self.L.pclean_set_row_shard(self.h, cls, begin, end))

    def nccl_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.L.pclean_nccl_init(self.h, buf, rank, world))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if lib().pclean_nccl_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        return buf.raw

    def decode(self, cell):
        tag = int(cell["tag"])
        if tag == VAL_STR:
            return self.string(int(cell["i"]))
        if tag == VAL_KEY:
            return int(cell["d"])
        if tag == 3:
            return float(cell["d"])
        return None


def load_trace_from_snapshot(engine: Engine, ir: FlatIR, model: M.PCleanModel, obs_cls_name: str, snapshot: dict):
    """Install a trace given as {class name: (keys, cells[nv, n], strength, discount)} plus
    {'assignment': {fk_vertex0: keys[n_rows]}} and {'params': {slot: values}}."""
    for name, (keys, cells, s, d) in snapshot["tables"].items():
        engine.load_table(ir.class_index[name], keys, cells, s, d)
    fks = sorted(snapshot["assignment"].keys())
    engine.load_assignment(ir.class_index[obs_cls_name], fks, np.stack([snapshot["assignment"][f] for f in fks]))
    for slot, vals in snapshot.get("params", {}).items():
        engine.set_param(slot, list(vals))
    for v, cells in snapshot.get("rowcells", {}).items():
        engine.load_row_cells(ir.class_index[obs_cls_name], v, cells)
