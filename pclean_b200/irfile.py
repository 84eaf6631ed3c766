"""On-disk form of the flat model IR (`pclean_model_ir`, include/pclean_b200.h) and of
an observed dataset (`pclean_observations`), so that a Julia host (julia/PCleanB200.jl writes the
same file), the Python harness and the C ABI (`pclean_load_model_file`) share inputs.

Layout (little endian):
    char     magic[8]  = "PCLIRv1\\n"
    uint32   n_entries
    entry := uint32 name_len, char name[name_len], uint32 dtype, uint64 count,
             zero padding to the next multiple of 8 (file offset), count * itemsize payload bytes
    dtype: 0 int32 | 1 int64 | 2 float64 | 3 uint32 | 4 pclean_value (int32 tag, int32 i, float64 d) | 5 uint8
Every pointer field of `pclean_model_ir` is an entry named like the field; every scalar field an int32
entry of count 1.  Observations: "obs.cls", "obs.n_rows" (int64), "obs.n_cols", "obs.vertex_of_col",
"obs.cells" ([n_cols][n_rows] column-major).  "class_names" / "obs.columns" (uint8, newline-joined
UTF-8) are conveniences for harnesses; the engine ignores them.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Dict, List, Optional

import numpy as np

from .lowering import FlatIR, ModelIR, Observations, VALUE_DTYPE, Value

MAGIC = b"PCLIRv1\n"
_DTYPES = {0: np.dtype(np.int32), 1: np.dtype(np.int64), 2: np.dtype(np.float64), 3: np.dtype(np.uint32), 4: VALUE_DTYPE, 5: np.dtype(np.uint8)}
_CODES = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float64): 2, np.dtype(np.uint32): 3, VALUE_DTYPE: 4, np.dtype(np.uint8): 5}
_SCALARS = ["n_classes", "n_vertices", "n_blocks", "n_paths", "n_funcs", "n_params", "n_param_slots", "n_lists", "n_xforms", "n_strings"]


def _entries_of(ir: FlatIR) -> Dict[str, np.ndarray]:
    c = ir.as_ctypes()
    out: Dict[str, np.ndarray] = {}
    for name in _SCALARS:
        out[name] = np.asarray([getattr(c, name)], dtype=np.int32)
    for name, arr in ir._arrays.items():
        out[name] = np.ascontiguousarray(arr)
    out["class_names"] = np.frombuffer("\n".join(ir.model.class_order).encode("utf-8"), dtype=np.uint8)
    return out


def save_ir(path: str, ir: FlatIR, obs: Optional[Observations] = None) -> None:
    entries = _entries_of(ir)
    if obs is not None:
        voc, cells = obs._keep
        entries["obs.cls"] = np.asarray([obs.cls], dtype=np.int32)
        entries["obs.n_rows"] = np.asarray([obs.n_rows], dtype=np.int64)
        entries["obs.n_cols"] = np.asarray([obs.n_cols], dtype=np.int32)
        entries["obs.vertex_of_col"] = np.ascontiguousarray(voc, dtype=np.int32)
        entries["obs.cells"] = np.ascontiguousarray(cells).reshape(-1)
        if getattr(obs, "columns", None):
            entries["obs.columns"] = np.frombuffer("\n".join(obs.columns).encode("utf-8"), dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(entries)))
        for name, arr in entries.items():
            nb = name.encode("utf-8")
            f.write(struct.pack("<I", len(nb))); f.write(nb)
            f.write(struct.pack("<IQ", _CODES[arr.dtype], arr.size))
            pad = (-f.tell()) % 8
            f.write(b"\0" * pad)
            f.write(arr.tobytes())


class LoadedIR:
    """A model IR (and optionally a dataset) read back from a file: quacks like FlatIR for `Engine`."""

    def __init__(self, entries: Dict[str, np.ndarray]):
        self.entries = entries
        self.class_names: List[str] = bytes(entries["class_names"]).decode("utf-8").split("\n") if "class_names" in entries else []
        self.class_index = {c: k for k, c in enumerate(self.class_names)}
        off, cp = entries["str_off"], entries["str_cp"]
        self.strings = ["".join(chr(int(x)) for x in cp[off[i]:off[i + 1]]) for i in range(int(entries["n_strings"][0]))]
        self._ctypes = None

    def as_ctypes(self) -> ModelIR:
        ir = ModelIR()
        for name in _SCALARS:
            setattr(ir, name, int(self.entries[name][0]))
        for name, ctype in ModelIR._fields_:
            if name in _SCALARS:
                continue
            arr = self.entries[name]
            if ctype == C.POINTER(Value):
                setattr(ir, name, C.cast(arr.ctypes.data, C.POINTER(Value)))
            else:
                setattr(ir, name, arr.ctypes.data_as(ctype))
        self._ctypes = ir
        return ir

    def observations(self) -> Optional[Observations]:
        e = self.entries
        if "obs.cells" not in e:
            return None
        voc = np.ascontiguousarray(e["obs.vertex_of_col"], dtype=np.int32)
        cells = np.ascontiguousarray(e["obs.cells"])
        obs = Observations(int(e["obs.cls"][0]), int(e["obs.n_rows"][0]), int(e["obs.n_cols"][0]),
                           voc.ctypes.data_as(C.POINTER(C.c_int32)), C.cast(cells.ctypes.data, C.POINTER(Value)))
        obs._keep = (voc, cells)
        obs.columns = bytes(e["obs.columns"]).decode("utf-8").split("\n") if "obs.columns" in e else []
        return obs


def load_ir(path: str) -> LoadedIR:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a PCLIRv1 file")
    (n,) = struct.unpack_from("<I", buf, 8)
    pos = 12
    entries: Dict[str, np.ndarray] = {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, pos); pos += 4
        name = buf[pos:pos + ln].decode("utf-8"); pos += ln
        code, count = struct.unpack_from("<IQ", buf, pos); pos += 12
        pos += (-pos) % 8
        dt = _DTYPES[code]
        entries[name] = np.frombuffer(buf, dtype=dt, count=count, offset=pos).copy()
        pos += count * dt.itemsize
    return LoadedIR(entries)
