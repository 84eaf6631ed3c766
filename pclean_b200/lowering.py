"""Flatten a PCleanModel + Query + data into the POD arrays of `include/pclean_b200.h`.

This is the job of the (thin) host shim: walk the model IR (`src/model/model.jl:87-188`),
dictionary-encode strings, and tabulate each JuliaNode closure over the finite product of
its arguments' supports (SURVEY §7 "Hard parts") — the device cannot run closures.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .host_fixture import model as M

# value tags (pclean_b200.h)
VAL_ABSENT, VAL_MISSING, VAL_STR, VAL_REAL, VAL_INT, VAL_LIST, VAL_XFORM, VAL_PARAM, VAL_IPARAM, VAL_KEY, VAL_DUMMY = range(11)
NODE_JULIA, NODE_CHOICE, NODE_PARAM, NODE_FK = range(4)
WRAP_NONE, WRAP_SUBMODEL, WRAP_EXTERNAL = range(3)
FUNC_CONST, FUNC_TABLE, FUNC_ROUND_BACKWARD, FUNC_JOIN = range(4)

VALUE_DTYPE = np.dtype([("tag", np.int32), ("i", np.int32), ("d", np.float64)], align=True)
assert VALUE_DTYPE.itemsize == 16


class Value(C.Structure):
    _fields_ = [("tag", C.c_int32), ("i", C.c_int32), ("d", C.c_double)]


def _ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class ModelIR(C.Structure):
    _fields_ = [
        ("n_classes", C.c_int32), ("class_voff", C.POINTER(C.c_int32)),
        ("py_strength", C.POINTER(C.c_double)), ("py_discount", C.POINTER(C.c_double)),
        ("n_vertices", C.c_int32),
        ("v_kind", C.POINTER(C.c_int32)), ("v_wrap", C.POINTER(C.c_int32)),
        ("v_wrap_off", C.POINTER(C.c_int32)), ("wrap_fk", C.POINTER(C.c_int32)), ("wrap_subid", C.POINTER(C.c_int32)),
        ("v_dist", C.POINTER(C.c_int32)), ("v_args_off", C.POINTER(C.c_int32)), ("v_args", C.POINTER(C.c_int32)),
        ("v_func", C.POINTER(C.c_int32)), ("v_target", C.POINTER(C.c_int32)),
        ("v_vmap_off", C.POINTER(C.c_int32)), ("v_vmap", C.POINTER(C.c_int32)),
        ("v_param", C.POINTER(C.c_int32)), ("v_path", C.POINTER(C.c_int32)), ("v_extv", C.POINTER(C.c_int32)),
        ("class_block_off", C.POINTER(C.c_int32)), ("n_blocks", C.c_int32),
        ("block_voff", C.POINTER(C.c_int32)), ("block_v", C.POINTER(C.c_int32)),
        ("plan_off", C.POINTER(C.c_int32)), ("plan_vertex", C.POINTER(C.c_int32)), ("plan_nchild", C.POINTER(C.c_int32)),
        ("class_hash_off", C.POINTER(C.c_int32)), ("hash_v", C.POINTER(C.c_int32)),
        ("n_paths", C.c_int32), ("path_target", C.POINTER(C.c_int32)),
        ("path_len_off", C.POINTER(C.c_int32)), ("path_class", C.POINTER(C.c_int32)), ("path_vertex", C.POINTER(C.c_int32)),
        ("path_vmap_off", C.POINTER(C.c_int32)), ("path_vmap", C.POINTER(C.c_int32)),
        ("n_funcs", C.c_int32), ("func_kind", C.POINTER(C.c_int32)), ("func_const", C.POINTER(Value)),
        ("func_keyarg_off", C.POINTER(C.c_int32)), ("func_keyargs", C.POINTER(C.c_int32)),
        ("func_tab_off", C.POINTER(C.c_int32)), ("tab_keys", C.POINTER(C.c_int32)),
        ("tab_key_off", C.POINTER(C.c_int64)), ("tab_vals", C.POINTER(Value)),
        ("n_params", C.c_int32), ("param_kind", C.POINTER(C.c_int32)), ("param_indexed", C.POINTER(C.c_int32)),
        ("param_prior0", C.POINTER(C.c_double)), ("param_prior1", C.POINTER(C.c_double)),
        ("n_param_slots", C.c_int32), ("slot_param", C.POINTER(C.c_int32)),
        ("n_lists", C.c_int32), ("list_off", C.POINTER(C.c_int64)), ("list_vals", C.POINTER(Value)),
        ("n_xforms", C.c_int32), ("xform_scale", C.POINTER(C.c_double)),
        ("n_strings", C.c_int32), ("str_off", C.POINTER(C.c_int64)), ("str_cp", C.POINTER(C.c_uint32)),
        ("lm_unigram", C.POINTER(C.c_double)), ("lm_bigram", C.POINTER(C.c_double)),
    ]


class Observations(C.Structure):
    _fields_ = [("cls", C.c_int32), ("n_rows", C.c_int64), ("n_cols", C.c_int32),
                ("vertex_of_col", C.POINTER(C.c_int32)), ("cells", C.POINTER(Value))]


class Config(C.Structure):
    _fields_ = [("num_iters", C.c_int32), ("num_particles", C.c_int32), ("use_dd_proposals", C.c_int32),
                ("use_lo_sweeps", C.c_int32), ("use_mh_instead_of_pg", C.c_int32),
                ("rejuv_frequency", C.c_int32), ("reporting_frequency", C.c_int32)]

    @classmethod
    def from_config(cls, cfg: M.InferenceConfig) -> "Config":
        return cls(cfg.num_iters, cfg.num_particles, int(cfg.use_dd_proposals), int(cfg.use_lo_sweeps),
                   int(cfg.use_mh_instead_of_pg), cfg.rejuv_frequency, cfg.reporting_frequency)


_LM_DIR = os.path.join(os.path.dirname(__file__), "lmparams")


def load_language_model() -> Tuple[np.ndarray, np.ndarray]:
    """28-symbol unigram / bigram tables of StringPrior (string_prior.jl:6-9).
    bigram[next, prev]: the reference indexes `english_letter_transitions[:, prev]`."""
    uni = np.loadtxt(os.path.join(_LM_DIR, "letter_probabilities.csv"), delimiter=",").reshape(-1)
    big = np.loadtxt(os.path.join(_LM_DIR, "letter_transition_matrix.csv"), delimiter=",")
    assert uni.shape == (28,) and big.shape == (28, 28)
    return np.ascontiguousarray(uni, dtype=np.float64), np.ascontiguousarray(big, dtype=np.float64)


class FlatIR:
    """Owns the numpy arrays and the ctypes view passed across the C ABI."""

    def __init__(self, model: M.PCleanModel, datasets: List[M.ObservedDataset] = ()):
        self.model = model
        self.class_index = {c: k for k, c in enumerate(model.class_order)}
        self.strings: List[str] = []
        self.string_id: Dict[str, int] = {}
        self.lists: List[Tuple] = []
        self.list_id: Dict[Tuple, int] = {}
        self.xforms: List[float] = []
        self.xform_id: Dict[float, int] = {}
        # parameter specs / slots
        self.param_spec: Dict[Tuple[str, int], int] = {}
        self.param_kind: List[int] = []
        self.param_indexed: List[int] = []
        self.param_prior: List[Tuple[float, float]] = []
        self.slot_param: List[int] = []
        self.slot_id: Dict[Tuple[int, Any], int] = {}
        self.slot_key: List[Any] = []
        # functions
        self.func_id: Dict[int, int] = {}
        self.func_kind: List[int] = []
        self.func_const: List[Tuple[int, int, float]] = []
        self.func_keyargs: List[List[int]] = []
        self.func_tab: List[List[Tuple[Tuple[int, ...], Tuple[int, int, float]]]] = []
        self._support_cache: Dict[Tuple[str, int], Optional[List]] = {}
        self._data_support: Dict[Tuple[str, int], List] = {}
        for ds in datasets:
            self._register_data_support(ds)
        self._build()

    # ----------------------------------------------------------------- interning
    def intern_string(self, s: str) -> int:
        k = self.string_id.get(s)
        if k is None:
            k = len(self.strings)
            self.strings.append(s)
            self.string_id[s] = k
        return k

    def encode(self, v) -> Tuple[int, int, float]:
        if v is None:
            return (VAL_MISSING, 0, 0.0)
        if isinstance(v, str):
            return (VAL_STR, self.intern_string(v), 0.0)
        if isinstance(v, bool):
            return (VAL_INT, int(v), 0.0)
        if isinstance(v, int):
            return (VAL_INT, v, 0.0)
        if isinstance(v, float):
            return (VAL_REAL, 0, v)
        if isinstance(v, M.Transformation):
            k = self.xform_id.get(v.scale)
            if k is None:
                k = len(self.xforms)
                self.xforms.append(v.scale)
                self.xform_id[v.scale] = k
            return (VAL_XFORM, k, 0.0)
        if isinstance(v, M.ParamSlot):
            spec = self.param_spec[(v.class_name, v.vertex)]
            return (VAL_PARAM, self._slot(spec, v.key), 0.0)
        if isinstance(v, M.ParamHandle):
            spec = self.param_spec[(v.class_name, v.vertex)]
            if self.param_indexed[spec]:
                return (VAL_IPARAM, spec, 0.0)
            return (VAL_PARAM, self._slot(spec, None), 0.0)
        if isinstance(v, (list, tuple)):
            enc = tuple(self.encode(x) for x in v)
            k = self.list_id.get(enc)
            if k is None:
                k = len(self.lists)
                self.lists.append(enc)
                self.list_id[enc] = k
            return (VAL_LIST, k, 0.0)
        raise TypeError(f"cannot encode value of type {type(v)}: {v!r}")

    def _slot(self, spec: int, key) -> int:
        k = self.slot_id.get((spec, key))
        if k is None:
            k = len(self.slot_param)
            self.slot_param.append(spec)
            self.slot_key.append(key)
            self.slot_id[(spec, key)] = k
        return k

    # ----------------------------------------------------------------- supports
    def _register_data_support(self, ds: M.ObservedDataset):
        q = ds.query
        for col, v in list(q.obsmap.items()) + list(q.cleanmap.items()):
            if col not in ds.data:
                continue
            vals = []
            seen = set()
            for x in ds.data[col]:
                if x is not None and x not in seen:
                    seen.add(x)
                    vals.append(x)
            self._register(q.cls, v, vals)

    def _register(self, cls: str, v: int, vals: List):
        cur = self._data_support.setdefault((cls, v), [])
        have = set(cur)
        for x in vals:
            if x not in have:
                have.add(x)
                cur.append(x)
        node = self.model.classes[cls].node(v)
        if isinstance(node, M.SubmodelNode):
            fk = M.strip_subnodes(self.model.classes[cls].node(node.foreign_key_node_id))
            self._register(fk.target_class, node.subnode_id, vals)

    def support(self, cls: str, v: int) -> Optional[List]:
        """Finite set of values vertex v can take on a scoring path (None = unknown/continuous)."""
        key = (cls, v)
        if key in self._support_cache:
            return self._support_cache[key]
        self._support_cache[key] = None   # cycle guard
        cm = self.model.classes[cls]
        node = cm.node(v)
        out: Optional[List] = None
        if isinstance(node, M.SubmodelNode):
            fk = M.strip_subnodes(cm.node(node.foreign_key_node_id))
            out = self.support(fk.target_class, node.subnode_id)
        elif isinstance(node, M.ParameterNode):
            out = [M.ParamHandle(cls, v)]
        elif isinstance(node, M.JuliaNode):
            out = self._julia_outputs(cls, node)
        elif isinstance(node, M.RandomChoiceNode):
            d = node.dist
            if d in (M.DIST_CHOOSE_PROPORTIONALLY, M.DIST_CHOOSE_UNIFORMLY):
                out = self._flatten_lists(self.support(cls, node.arg_node_ids[0]))
            elif d == M.DIST_STRING_PRIOR:
                out = self._flatten_lists(self.support(cls, node.arg_node_ids[2]))
            elif d == M.DIST_TIME_PRIOR:
                out = self._flatten_lists(self.support(cls, node.arg_node_ids[0]))
            data = self._data_support.get(key)
            if data:
                out = _union(out or [], data)
        if out is None and key in self._data_support:
            out = list(self._data_support[key])
        self._support_cache[key] = out
        return out

    @staticmethod
    def _flatten_lists(lists: Optional[List]) -> Optional[List]:
        if lists is None:
            return None
        out: List = []
        seen = set()
        for lst in lists:
            for x in lst:
                if x not in seen:
                    seen.add(x)
                    out.append(x)
        return out

    def _julia_outputs(self, cls: str, node: M.JuliaNode) -> Optional[List]:
        if node.builtin is not None and node.builtin[0] == "round_backward":
            return None
        sups = []
        for a in node.arg_node_ids:
            s = self.support(cls, a)
            if s is None:
                return None
            sups.append(s)
        out = []
        seen = set()
        total = 1
        for s in sups:
            total *= max(1, len(s))
        if total > 5_000_000:
            return None
        for combo in itertools.product(*sups):
            try:
                r = node.f(*combo)
            except (KeyError, IndexError):
                continue
            h = _hashable(r)
            if h not in seen:
                seen.add(h)
                out.append(r)
        return out

    # ----------------------------------------------------------------- functions
    def _function(self, cls: str, node: M.JuliaNode) -> int:
        fid = self.func_id.get(id(node.f))
        if fid is not None:
            return fid
        fid = len(self.func_kind)
        self.func_id[id(node.f)] = fid
        self.func_kind.append(FUNC_CONST)
        self.func_const.append((VAL_ABSENT, 0, 0.0))
        self.func_keyargs.append([])
        self.func_tab.append([])
        if node.builtin is not None:
            if node.builtin[0] == "round_backward":
                self.func_kind[fid] = FUNC_ROUND_BACKWARD
            elif node.builtin[0] == "join":
                self.func_kind[fid] = FUNC_JOIN
                self.func_const[fid] = self.encode(node.builtin[1])
            else:
                raise ValueError(node.builtin)
            return fid
        if not node.arg_node_ids:
            self.func_const[fid] = self.encode(node.f())
            return fid
        self.func_kind[fid] = FUNC_TABLE
        sups = []
        keypos = []
        for pos, a in enumerate(node.arg_node_ids):
            s = self.support(cls, a)
            if s is None:
                raise ValueError(f"JuliaNode in class {cls} has an argument (vertex {a}) with no finite support; "
                                 "supply a builtin")
            sups.append(s)
            if not (len(s) == 1 and isinstance(s[0], M.ParamHandle)):
                keypos.append(pos)
        self.func_keyargs[fid] = keypos
        entries = []
        for combo in itertools.product(*sups):
            try:
                r = node.f(*combo)
            except (KeyError, IndexError):
                continue
            key = tuple(self._key_code(combo[p]) for p in keypos)
            entries.append((key, self.encode(r)))
        self.func_tab[fid] = entries
        return fid

    def _key_code(self, v) -> int:
        tag, i, _ = self.encode(v)
        if tag not in (VAL_STR, VAL_INT, VAL_LIST, VAL_XFORM, VAL_PARAM, VAL_IPARAM):
            raise TypeError(f"JuliaNode key argument must be discrete, got {v!r}")
        return i

    # ----------------------------------------------------------------- build
    def _build(self):
        model = self.model
        order = model.class_order
        # parameter specs first (class order, vertex order) so slots of basic params are stable
        for cls in order:
            cm = model.classes[cls]
            for v, node in enumerate(cm.nodes, 1):
                if isinstance(node, M.ParameterNode):
                    spec = len(self.param_kind)
                    self.param_spec[(cls, v)] = spec
                    self.param_kind.append(node.kind)
                    self.param_indexed.append(int(node.indexed))
                    self.param_prior.append(node.prior)
                    if not node.indexed:
                        self._slot(spec, None)

        class_voff = [0]
        v_kind, v_wrap, v_dist, v_func, v_target, v_param, v_path, v_extv = [], [], [], [], [], [], [], []
        v_wrap_off, wrap_fk, wrap_subid = [0], [], []
        v_args_off, v_args = [0], []
        v_vmap_off, v_vmap = [0], []
        class_block_off, block_voff, block_v = [0], [0], []
        plan_off, plan_vertex, plan_nchild = [0], [], []
        class_hash_off, hash_v = [0], []
        path_ids: Dict[Tuple, int] = {}
        path_target, path_len_off, path_class, path_vertex = [], [0], [], []
        path_vmap_off, path_vmap = [0], []

        # paths: enumerate per class in class order
        for cls in order:
            cm = model.classes[cls]
            n_normal = sum(1 for n in cm.nodes if not isinstance(n, M.ExternalLikelihoodNode))
            for path, vmap in cm.incoming_references.items():
                path_ids[(cls, path)] = len(path_target)
                path_target.append(self.class_index[cls])
                for (pc, pv) in path:
                    path_class.append(self.class_index[pc])
                    path_vertex.append(pv - 1)
                path_len_off.append(len(path_class))
                dense = [-1] * n_normal
                for i, j in vmap.items():
                    dense[i - 1] = j - 1
                path_vmap.extend(dense)
                path_vmap_off.append(len(path_vmap))

        for cls in order:
            cm = model.classes[cls]
            for v, node in enumerate(cm.nodes, 1):
                wrap = WRAP_NONE
                base = node
                ext_path, ext_v = -1, -1
                func_cls = cls
                if isinstance(node, M.ExternalLikelihoodNode):
                    wrap = WRAP_EXTERNAL
                    base = node.external_node
                    ext_path = path_ids[(cls, node.path)]
                    ext_v = node.external_node_id - 1
                    func_cls = node.path[-1][0]
                elif isinstance(node, M.SubmodelNode):
                    wrap = WRAP_SUBMODEL
                    while isinstance(base, M.SubmodelNode):
                        wrap_fk.append(base.foreign_key_node_id - 1)
                        wrap_subid.append(base.subnode_id - 1)
                        base = base.subnode
                v_wrap_off.append(len(wrap_fk))
                v_wrap.append(wrap)
                v_path.append(ext_path)
                v_extv.append(ext_v)
                kind, dist, func, target, param = -1, -1, -1, -1, -1
                args: List[int] = []
                vmap_dense: List[int] = []
                if isinstance(base, M.JuliaNode):
                    kind = NODE_JULIA
                    args = [a - 1 for a in base.arg_node_ids]
                    func = self._function_for(func_cls, base)
                elif isinstance(base, M.RandomChoiceNode):
                    kind = NODE_CHOICE
                    dist = base.dist
                    args = [a - 1 for a in base.arg_node_ids]
                elif isinstance(base, M.ParameterNode):
                    kind = NODE_PARAM
                    param = self._param_spec_of(base)
                elif isinstance(base, M.ForeignKeyNode):
                    kind = NODE_FK
                    target = self.class_index[base.target_class]
                    n_t = len(base.vmap)
                    vmap_dense = [base.vmap[i] - 1 for i in range(1, n_t + 1)]
                else:
                    raise TypeError(base)
                v_kind.append(kind); v_dist.append(dist); v_func.append(func)
                v_target.append(target); v_param.append(param)
                v_args.extend(args); v_args_off.append(len(v_args))
                v_vmap.extend(vmap_dense); v_vmap_off.append(len(v_vmap))
            class_voff.append(len(v_kind))
            for block, plan in zip(cm.blocks, cm.plans):
                block_v.extend(x - 1 for x in block)
                block_voff.append(len(block_v))
                _flatten_plan(plan, plan_vertex, plan_nchild)
                plan_off.append(len(plan_vertex))
            class_block_off.append(len(block_voff) - 1)
            hash_v.extend(h - 1 for h in cm.hash_keys)
            class_hash_off.append(len(hash_v))

        self.n_classes = len(order)
        i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
        i64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int64))
        f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        a: Dict[str, np.ndarray] = {}
        a["class_voff"] = i32(class_voff)
        a["py_strength"] = f64([model.classes[c].initial_pitman_yor_params[0] for c in order])
        a["py_discount"] = f64([model.classes[c].initial_pitman_yor_params[1] for c in order])
        for name, val in [("v_kind", v_kind), ("v_wrap", v_wrap), ("v_wrap_off", v_wrap_off), ("wrap_fk", wrap_fk),
                          ("wrap_subid", wrap_subid), ("v_dist", v_dist), ("v_args_off", v_args_off),
                          ("v_args", v_args), ("v_func", v_func), ("v_target", v_target),
                          ("v_vmap_off", v_vmap_off), ("v_vmap", v_vmap), ("v_param", v_param),
                          ("v_path", v_path), ("v_extv", v_extv), ("class_block_off", class_block_off),
                          ("block_voff", block_voff), ("block_v", block_v), ("plan_off", plan_off),
                          ("plan_vertex", plan_vertex), ("plan_nchild", plan_nchild),
                          ("class_hash_off", class_hash_off), ("hash_v", hash_v),
                          ("path_target", path_target), ("path_len_off", path_len_off),
                          ("path_class", path_class), ("path_vertex", path_vertex),
                          ("path_vmap_off", path_vmap_off), ("path_vmap", path_vmap)]:
            a[name] = i32(val)
        self.n_vertices = len(v_kind)
        self.n_blocks = len(block_voff) - 1
        self.n_paths = len(path_target)
        self.path_ids = path_ids
        self._arrays = a
        self._finalize_tables()

    def _param_spec_of(self, node: M.ParameterNode) -> int:
        for (cls, v), spec in self.param_spec.items():
            if self.model.classes[cls].node(v) is node:
                return spec
        raise KeyError("parameter node not registered")

    def _function_for(self, cls: str, node: M.JuliaNode) -> int:
        fid = self.func_id.get(id(node.f))
        if fid is not None:
            return fid
        # tabulate in the class where the closure was declared: find the un-shifted original
        for c in self.model.class_order:
            for n in self.model.classes[c].nodes:
                if isinstance(n, M.JuliaNode) and n.f is node.f:
                    return self._function(c, n)
        return self._function(cls, node)

    def _finalize_tables(self):
        a = self._arrays
        i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
        i64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int64))
        f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))

        def vals(lst):
            arr = np.zeros(max(1, len(lst)), dtype=VALUE_DTYPE)
            for k, (t, i, d) in enumerate(lst):
                arr[k] = (t, i, d)
            return arr

        a["func_kind"] = i32(self.func_kind)
        a["func_const"] = vals(self.func_const)
        ko, ka = [0], []
        to, tk, tko, tv = [0], [], [0], []
        for fid in range(len(self.func_kind)):
            ka.extend(self.func_keyargs[fid]); ko.append(len(ka))
            for key, val in self.func_tab[fid]:
                tk.extend(key); tko.append(len(tk)); tv.append(val)
            to.append(len(tv))
        a["func_keyarg_off"] = i32(ko); a["func_keyargs"] = i32(ka if ka else [0])
        a["func_tab_off"] = i32(to); a["tab_keys"] = i32(tk if tk else [0])
        a["tab_key_off"] = i64(tko); a["tab_vals"] = vals(tv)
        a["param_kind"] = i32(self.param_kind if self.param_kind else [0])
        a["param_indexed"] = i32(self.param_indexed if self.param_indexed else [0])
        a["param_prior0"] = f64([p[0] for p in self.param_prior] or [0.0])
        a["param_prior1"] = f64([p[1] for p in self.param_prior] or [0.0])
        a["slot_param"] = i32(self.slot_param if self.slot_param else [0])
        lo, lv = [0], []
        for lst in self.lists:
            lv.extend(lst); lo.append(len(lv))
        a["list_off"] = i64(lo); a["list_vals"] = vals(lv)
        a["xform_scale"] = f64(self.xforms if self.xforms else [1.0])
        so, cp = [0], []
        for s in self.strings:
            cp.extend(ord(ch) for ch in s); so.append(len(cp))
        a["str_off"] = i64(so); a["str_cp"] = np.ascontiguousarray(np.asarray(cp if cp else [0], dtype=np.uint32))
        uni, big = load_language_model()
        a["lm_unigram"], a["lm_bigram"] = uni, big
        self._ctypes = None

    # ----------------------------------------------------------------- ctypes view
    def refresh(self):
        """Re-emit dictionary/list/function arrays after new values were interned
        (e.g. by encode_observations)."""
        self._finalize_tables()

    def as_ctypes(self) -> ModelIR:
        a = self._arrays
        ir = ModelIR()
        ir.n_classes = self.n_classes
        ir.n_vertices = self.n_vertices
        ir.n_blocks = self.n_blocks
        ir.n_paths = self.n_paths
        ir.n_funcs = len(self.func_kind)
        ir.n_params = len(self.param_kind)
        ir.n_param_slots = len(self.slot_param)
        ir.n_lists = len(self.lists)
        ir.n_xforms = len(self.xforms)
        ir.n_strings = len(self.strings)
        for name, ctype in ModelIR._fields_:
            if name in a:
                arr = a[name]
                if ctype == C.POINTER(Value):
                    setattr(ir, name, C.cast(arr.ctypes.data, C.POINTER(Value)))
                else:
                    setattr(ir, name, arr.ctypes.data_as(ctype))
        self._ctypes = ir
        return ir

    # ----------------------------------------------------------------- observations
    def encode_observations(self, ds: M.ObservedDataset):
        """inference.jl:17-33 — build the per-row observation cells (column-major)."""
        q = ds.query
        cm = self.model.classes[q.cls]
        cols = [c for c in ds.data.keys() if c in q.obsmap]
        n = len(next(iter(ds.data.values())))
        cells = np.zeros((len(cols), n), dtype=VALUE_DTYPE)
        vertex_of_col = []
        for ci, col in enumerate(cols):
            node_id = q.obsmap[col]
            vertex_of_col.append(node_id - 1)
            base = M.strip_subnodes(cm.node(node_id))
            explicit_missing = (node_id != q.cleanmap[col] and isinstance(base, M.RandomChoiceNode)
                                and base.dist in M.SUPPORTS_MISSING)
            data = ds.data[col]
            tags = np.zeros(n, dtype=np.int32)
            ii = np.zeros(n, dtype=np.int32)
            dd = np.zeros(n, dtype=np.float64)
            cache: Dict[Any, Tuple[int, int, float]] = {}
            for r, x in enumerate(data):
                if x is None:
                    tags[r] = VAL_MISSING if explicit_missing else VAL_ABSENT
                    continue
                e = cache.get(x)
                if e is None:
                    e = self.encode(x)
                    cache[x] = e
                tags[r], ii[r], dd[r] = e
            cells[ci]["tag"] = tags; cells[ci]["i"] = ii; cells[ci]["d"] = dd
        self.refresh()
        voc = np.ascontiguousarray(np.asarray(vertex_of_col, dtype=np.int32))
        cells = np.ascontiguousarray(cells)
        obs = Observations(self.class_index[q.cls], n, len(cols),
                           voc.ctypes.data_as(C.POINTER(C.c_int32)),
                           C.cast(cells.ctypes.data, C.POINTER(Value)))
        obs._keep = (voc, cells)
        obs.columns = cols
        return obs


def _flatten_plan(plan: M.Plan, out_vertex: List[int], out_nchild: List[int]):
    """Preorder (vertex, n_children); the forest root is implicit: a leading pseudo entry
    (-1, n_roots) so a block's plan is self-delimiting."""
    def rec(p: M.Plan):
        for s in p.steps:
            out_vertex.append(s.idx - 1)
            out_nchild.append(len(s.rest.steps))
            rec(s.rest)
    out_vertex.append(-1)
    out_nchild.append(len(plan.steps))
    rec(plan)


def _hashable(x):
    if isinstance(x, list):
        return tuple(_hashable(y) for y in x)
    return x


def _union(a: List, b: List) -> List:
    seen = set(_hashable(x) for x in a)
    out = list(a)
    for x in b:
        h = _hashable(x)
        if h not in seen:
            seen.add(h)
            out.append(x)
    return out
