"""Model IR and builder — host-side mirror of the reference's model layer.

The Julia host keeps `PClean.@model/@class/@query` unchanged; this module plays that
host in Python (no Julia toolchain in the build image).  It reproduces the *builder
commands* the `@model` macro expands to (reference `src/dsl/syntax.jl:106-161`) and
the IR they construct (`src/model/model.jl:87-188`, `src/dsl/builder.jl`), with the
same vertex numbering, block lists, plans, vmaps and incoming-reference paths, so that
the flat IR handed to the C-ABI engine is the one a Julia shim would export.

Vertex ids are 1-based here exactly as in the reference; `lowering.py` converts to the
0-based flat arrays of `include/pclean_b200.h`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

# --------------------------------------------------------------------------------------
# Distribution descriptors (closed set, reference src/distributions/*.jl).  The numeric
# ids are the `PCLEAN_DIST_*` constants of include/pclean_b200.h.
# --------------------------------------------------------------------------------------
DIST_ADD_TYPOS = 0             # add_typos.jl
DIST_CHOOSE_PROPORTIONALLY = 1 # choose_proportionally.jl
DIST_CHOOSE_UNIFORMLY = 2      # choose_uniformly.jl
DIST_STRING_PRIOR = 3          # string_prior.jl
DIST_TIME_PRIOR = 4            # time_prior.jl
DIST_MAYBE_SWAP = 5            # maybe_swap.jl
DIST_TRANSFORMED_GAUSSIAN = 6  # transformed_gaussian.jl
DIST_UNMODELED = 7             # unmodeled.jl
DIST_ADD_NOISE = 8             # add_noise.jl

DIST_NAMES = {
    DIST_ADD_TYPOS: "AddTypos", DIST_CHOOSE_PROPORTIONALLY: "ChooseProportionally",
    DIST_CHOOSE_UNIFORMLY: "ChooseUniformly", DIST_STRING_PRIOR: "StringPrior",
    DIST_TIME_PRIOR: "TimePrior", DIST_MAYBE_SWAP: "MaybeSwap",
    DIST_TRANSFORMED_GAUSSIAN: "TransformedGaussian", DIST_UNMODELED: "Unmodeled",
    DIST_ADD_NOISE: "AddNoise",
}
# has_discrete_proposal(dist) — reference: add_typos.jl:5, choose_proportionally.jl:13,
# choose_uniformly.jl:12, string_prior.jl:11, time_prior.jl:5, maybe_swap.jl:30,
# transformed_gaussian.jl:11, unmodeled.jl:13, add_noise.jl:3
HAS_DISCRETE_PROPOSAL = {
    DIST_ADD_TYPOS: False, DIST_CHOOSE_PROPORTIONALLY: True, DIST_CHOOSE_UNIFORMLY: True,
    DIST_STRING_PRIOR: True, DIST_TIME_PRIOR: True, DIST_MAYBE_SWAP: False,
    DIST_TRANSFORMED_GAUSSIAN: False, DIST_UNMODELED: False, DIST_ADD_NOISE: False,
}
# supports_explicitly_missing_observations — add_typos.jl:7, maybe_swap.jl:3, unmodeled.jl:15
SUPPORTS_MISSING = {DIST_ADD_TYPOS, DIST_MAYBE_SWAP, DIST_UNMODELED}


class AddTypos: dist_id = DIST_ADD_TYPOS
class ChooseProportionally: dist_id = DIST_CHOOSE_PROPORTIONALLY
class ChooseUniformly: dist_id = DIST_CHOOSE_UNIFORMLY
class StringPrior: dist_id = DIST_STRING_PRIOR
class TimePrior: dist_id = DIST_TIME_PRIOR
class MaybeSwap: dist_id = DIST_MAYBE_SWAP
class TransformedGaussian: dist_id = DIST_TRANSFORMED_GAUSSIAN
class Unmodeled: dist_id = DIST_UNMODELED
class AddNoise: dist_id = DIST_ADD_NOISE


# Parameter kinds (PCLEAN_PARAM_*)
PARAM_PROPORTIONS = 0   # choose_proportionally.jl:31-74
PARAM_MEAN = 1          # add_noise.jl:21-82
PARAM_PROB = 2          # maybe_swap.jl:41-89


@dataclass(frozen=True)
class Transformation:
    """transformed_gaussian.jl:5-9.  The engine supports the affine family the shipped
    models use: backward(x) = x * scale, forward(x) = x / scale, |g'| = 1/scale."""
    scale: float = 1.0

    def forward(self, x): return x / self.scale
    def backward(self, x): return x * self.scale
    def deriv(self, x): return 1.0 / self.scale


@dataclass(frozen=True)
class ParamHandle:
    """What a JuliaNode sees when one of its arguments is a ParameterNode."""
    class_name: str
    vertex: int

    def __getitem__(self, key) -> "ParamSlot":   # IndexedParameter getindex, distributions.jl:50-55
        return ParamSlot(self.class_name, self.vertex, key)


@dataclass(frozen=True)
class ParamSlot:
    class_name: str
    vertex: int
    key: Any


# --------------------------------------------------------------------------------------
# Node types (model.jl:136-181)
# --------------------------------------------------------------------------------------
@dataclass
class JuliaNode:
    f: Callable
    arg_node_ids: List[int]
    # host-only hint: builtin understood natively by the engine, e.g. ("join", "_") or
    # ("round_backward",); None = tabulate the closure over its argument supports
    builtin: Any = None


@dataclass
class RandomChoiceNode:
    dist: int
    arg_node_ids: List[int]


@dataclass
class ParameterNode:
    kind: int
    indexed: bool
    prior: Tuple[float, float]


@dataclass
class ForeignKeyNode:
    target_class: str
    vmap: Dict[int, int]


@dataclass
class SubmodelNode:
    foreign_key_node_id: int
    subnode_id: int
    subnode: Any


@dataclass
class ExternalLikelihoodNode:
    path: Tuple[Tuple[str, int], ...]
    external_node_id: int
    external_node: Any


def strip_subnodes(node):
    while isinstance(node, SubmodelNode):
        node = node.subnode
    return node


@dataclass
class Step:
    idx: int
    rest: "Plan"


@dataclass
class Plan:
    steps: List[Step] = field(default_factory=list)


class DiGraph:
    """The tiny subset of LightGraphs.SimpleDiGraph the builder uses."""

    def __init__(self):
        self.out: List[List[int]] = [[]]   # 1-based; index 0 unused
        self.inn: List[List[int]] = [[]]

    def nv(self) -> int:
        return len(self.out) - 1

    def add_vertex(self) -> int:
        self.out.append([]); self.inn.append([])
        return self.nv()

    def add_edge(self, s: int, d: int) -> bool:
        if s < 1 or d < 1 or s > self.nv() or d > self.nv():
            return False       # LightGraphs silently rejects edges to missing vertices
        if d in self.out[s]:
            return False
        self.out[s].append(d); self.out[s].sort()
        self.inn[d].append(s); self.inn[d].sort()
        return True

    def outneighbors(self, v: int) -> List[int]:
        return list(self.out[v])

    def edges(self):
        for s in range(1, self.nv() + 1):
            for d in self.out[s]:
                yield s, d


@dataclass
class PCleanClass:
    graph: DiGraph = field(default_factory=DiGraph)
    nodes: List[Any] = field(default_factory=list)           # nodes[v-1]
    hash_keys: List[int] = field(default_factory=list)
    blocks: List[List[int]] = field(default_factory=list)
    plans: List[Plan] = field(default_factory=list)
    names: Dict[str, int] = field(default_factory=dict)
    incoming_references: Dict[Tuple[Tuple[str, int], ...], Dict[int, int]] = field(default_factory=dict)
    initial_pitman_yor_params: Tuple[float, float] = (1.0, 0.0)   # builder.jl:39

    def node(self, v: int):
        return self.nodes[v - 1]


@dataclass
class PCleanModel:
    classes: Dict[str, PCleanClass] = field(default_factory=dict)
    class_order: List[str] = field(default_factory=list)


# --------------------------------------------------------------------------------------
# Builder (builder.jl)
# --------------------------------------------------------------------------------------
OPEN, CLOSED = 0, 1


def copy_node(n, v: int):
    """builder.jl:115-120"""
    if isinstance(n, JuliaNode):
        return JuliaNode(n.f, [x + v for x in n.arg_node_ids], n.builtin)
    if isinstance(n, RandomChoiceNode):
        return RandomChoiceNode(n.dist, [x + v for x in n.arg_node_ids])
    if isinstance(n, ParameterNode):
        return n
    if isinstance(n, ForeignKeyNode):
        return ForeignKeyNode(n.target_class, {i: j + v for i, j in n.vmap.items()})
    if isinstance(n, SubmodelNode):
        return SubmodelNode(n.foreign_key_node_id + v, n.subnode_id, copy_node(n.subnode, v))
    raise TypeError(n)


Argument = Union[str, Tuple[List[str], Callable]]


class PCleanModelBuilder:
    """Mirror of `PCleanModelBuilder` (builder.jl:29-32).  Arguments are either a name
    (`"x"` or a dotted slot chain `"hosp.loc.city"`) or a `(names, function)` pair —
    exactly what `parse_expression` (syntax.jl:76-86) hands to the builder."""

    def __init__(self):
        self.model = PCleanModel()
        self.block_status = CLOSED
        self._gensym = 0

    # -- blocking (builder.jl:13-20)
    def begin_block(self, cls: str):
        self.model.classes[cls].blocks.append([])
        self.block_status = OPEN

    def end_block(self):
        self.block_status = CLOSED

    # -- builder.jl:37-41
    def add_new_class(self, cls: str):
        self.model.classes[cls] = PCleanClass()
        self.model.class_order.append(cls)

    # -- name resolution (builder.jl:63-99)
    def resolve_dot_expression(self, cls: str, expr: str) -> int:
        cm = self.model.classes[cls]
        if "." not in expr:
            return cm.names[expr]
        head, tail = expr.split(".", 1)
        fk = cm.node(cm.names[head])
        fk = strip_subnodes(fk)
        return fk.vmap[self.resolve_dot_expression(fk.target_class, tail)]

    def resolve_argument(self, cls: str, argument: Argument) -> int:
        cm = self.model.classes[cls]
        if isinstance(argument, str):
            return self.resolve_dot_expression(cls, argument)
        syms, f = argument[0], argument[1]
        builtin = argument[2] if len(argument) > 2 else None
        self._gensym += 1
        self.add_julia_node(cls, f"#arg{self._gensym}", syms, f, builtin)
        return cm.graph.nv()

    # -- builder.jl:104-106
    def add_guaranteed(self, cls: str, name: str):
        self.model.classes[cls].hash_keys.append(self.resolve_argument(cls, name))

    # -- builder.jl:123-175
    def add_foreign_key(self, source_class: str, name: str, target_class: str):
        sm = self.model.classes[source_class]
        tm = self.model.classes[target_class]
        v = sm.graph.add_vertex()
        sm.names[name] = v
        target_nodes = [n for n in tm.nodes if not isinstance(n, ExternalLikelihoodNode)]
        sm.nodes.append(ForeignKeyNode(target_class, {i: i + v for i in range(1, len(target_nodes) + 1)}))
        # builder.jl:139-149 — parent_nodes is computed after the push, so it includes the
        # new FK itself (self loop; edges to not-yet-existing vertices are rejected).
        parents = [(i, n) for i, n in enumerate(sm.nodes, 1)
                   if isinstance(n, ForeignKeyNode) and n.target_class == target_class]
        for i, pn in parents:
            sm.graph.add_edge(i, v)
            for sub in pn.vmap.values():
                sm.graph.add_edge(sub, v)
        for i, node in enumerate(target_nodes, 1):
            sm.graph.add_vertex()
            sm.nodes.append(SubmodelNode(v, i, copy_node(node, v)))
            sm.graph.add_edge(v, i + v)
        limit = sm.graph.nv()
        for s, d in tm.graph.edges():
            if s + v <= limit and d + v <= limit:
                sm.graph.add_edge(s + v, d + v)
        sampled = [v]
        for block in tm.blocks:
            sampled.extend(x + v for x in block if x + v <= limit)
        if self.block_status == OPEN:
            sm.blocks[-1].extend(sampled)
        else:
            sm.blocks.append(sampled)
            self.block_status = OPEN

    # -- builder.jl:182-202
    def add_basic_parameter(self, cls: str, name: str, kind: int, *args):
        cm = self.model.classes[cls]
        v = cm.graph.add_vertex()
        cm.names[name] = v
        cm.nodes.append(ParameterNode(kind, False, default_prior(kind, *args)))

    def add_indexed_parameter(self, cls: str, name: str, kind: int, *args):
        cm = self.model.classes[cls]
        v = cm.graph.add_vertex()
        cm.names[name] = v
        cm.nodes.append(ParameterNode(kind, True, default_prior(kind, *args)))

    # -- builder.jl:208-258
    def _place_in_block(self, cm: PCleanClass, v: int):
        if self.block_status == CLOSED:
            cm.blocks.append([v])
            self.block_status = OPEN
        else:
            cm.blocks[-1].append(v)

    def add_julia_node(self, cls: str, name: str, arguments: List[Argument], f: Callable, builtin=None):
        cm = self.model.classes[cls]
        arg_indices = [self.resolve_argument(cls, a) for a in arguments]
        v = cm.graph.add_vertex()
        cm.names[name] = v
        for a in arg_indices:
            cm.graph.add_edge(a, v)
        cm.nodes.append(JuliaNode(f, arg_indices, builtin))
        self._place_in_block(cm, v)

    def add_choice_node(self, cls: str, name: str, dist, arguments: List[Argument]):
        cm = self.model.classes[cls]
        arg_indices = [self.resolve_argument(cls, a) for a in arguments]
        v = cm.graph.add_vertex()
        cm.names[name] = v
        for a in arg_indices:
            cm.graph.add_edge(a, v)
        dist_id = dist if isinstance(dist, int) else dist.dist_id
        cm.nodes.append(RandomChoiceNode(dist_id, arg_indices))
        self._place_in_block(cm, v)

    # -- external nodes (builder.jl:264-350)
    def _add_external_nodes(self, model_node, node_id, block_id, path, tm: PCleanClass, sm: PCleanClass,
                            added: Dict[int, int], frm: Optional[int] = None):
        if isinstance(model_node, (ParameterNode, SubmodelNode)):
            return
        if node_id in added:
            if frm is not None:
                tm.graph.add_edge(frm, added[node_id])
            return
        nv = tm.graph.add_vertex()
        added[node_id] = nv
        if frm is not None:
            tm.graph.add_edge(frm, nv)
        tm.blocks[block_id].append(nv)
        tm.nodes.append(ExternalLikelihoodNode(path, node_id, model_node))
        if isinstance(model_node, JuliaNode):
            for nxt in sm.graph.outneighbors(node_id):
                self._add_external_nodes(sm.node(nxt), nxt, block_id, path, tm, sm, added, nv)

    def _process_reference(self, target_class: str, path, vmap: Dict[int, int]):
        source_class = path[-1][0]
        sm = self.model.classes[source_class]
        tm = self.model.classes[target_class]
        tm.incoming_references[path] = vmap
        added: Dict[int, int] = {}
        for block_idx in reversed(range(len(tm.blocks))):
            block = list(tm.blocks[block_idx])
            nodes_from_block = [(i, vmap[i]) for i in block
                                if not isinstance(tm.node(i), ExternalLikelihoodNode)]
            for target_node, source_node in nodes_from_block:
                for nxt in sm.graph.outneighbors(source_node):
                    self._add_external_nodes(sm.node(nxt), nxt, block_idx, path, tm, sm, added, target_node)
        for v, node in enumerate(list(tm.nodes), 1):
            if isinstance(node, ForeignKeyNode):
                new_path = ((target_class, v),) + path
                new_vmap = {i: vmap[j] for i, j in node.vmap.items()}
                self._process_reference(node.target_class, new_path, new_vmap)

    def _process_references(self, cls: str):
        cm = self.model.classes[cls]
        for v, node in enumerate(list(cm.nodes), 1):
            if isinstance(node, ForeignKeyNode):
                self._process_reference(node.target_class, ((cls, v),), node.vmap)

    # -- builder.jl:378-386
    def finish_class(self, cls: str):
        self._process_references(cls)
        self.block_status = CLOSED

    def finish_model(self) -> PCleanModel:
        for cm in self.model.classes.values():
            for block in cm.blocks:
                cm.plans.append(make_plan(cm.graph, block))
        return self.model


def default_prior(kind: int, *args) -> Tuple[float, float]:
    """default_prior: choose_proportionally.jl:37-40 (concentration), add_noise.jl:30-34
    (mean, std), maybe_swap.jl:53-55 (a, b)."""
    if kind == PARAM_PROPORTIONS:
        return (float(args[0]) if args else 1.0, 0.0)
    if kind == PARAM_MEAN:
        if len(args) == 1:
            return (float(args[0]), 0.5 * abs(float(args[0])))
        return (float(args[0]), float(args[1]))
    if kind == PARAM_PROB:
        if len(args) == 0:
            return (1.0, 3.0)
        if len(args) == 1:
            return (float(args[0]) * 4, (1 - float(args[0])) * 4)
        return (float(args[0]), float(args[1]))
    raise ValueError(kind)


def make_plan(graph: DiGraph, toposort: List[int]) -> Plan:
    """builder.jl:356-361: induced subgraph → weakly connected components (ordered by
    first appearance) → first vertex of each component is a Step; recurse."""
    if not toposort:
        return Plan([])
    inset = {v: k for k, v in enumerate(toposort)}
    parent = list(range(len(toposort)))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for v in toposort:
        for d in graph.out[v]:
            if d in inset and d != v:
                ra, rb = find(inset[v]), find(inset[d])
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    comps: Dict[int, List[int]] = {}
    order: List[int] = []
    for k, v in enumerate(toposort):
        r = find(k)
        if r not in comps:
            comps[r] = []
            order.append(r)
        comps[r].append(v)
    return Plan([Step(comps[r][0], make_plan(graph, comps[r][1:])) for r in order])


# --------------------------------------------------------------------------------------
# Query (query.jl)
# --------------------------------------------------------------------------------------
@dataclass
class Query:
    model: PCleanModel
    cls: str
    cleanmap: Dict[str, int] = field(default_factory=dict)
    obsmap: Dict[str, int] = field(default_factory=dict)


def make_query(builder_or_model, cls: str, rows: List[Tuple]) -> Query:
    """`@query Model.Class [col clean dirty; ...]` (query.jl:15-38).  A 2-tuple binds the
    same expression as clean and dirty."""
    model = builder_or_model.model if isinstance(builder_or_model, PCleanModelBuilder) else builder_or_model
    b = PCleanModelBuilder()
    b.model = model
    q = Query(model, cls)
    for row in rows:
        if len(row) == 2:
            col, clean = row
            dirty = clean
        else:
            col, clean, dirty = row
        q.cleanmap[col] = b.resolve_dot_expression(cls, clean)
        q.obsmap[col] = b.resolve_dot_expression(cls, dirty)
    return q


@dataclass
class ObservedDataset:
    query: Query
    data: Any       # dict column -> list of values (None = missing)


@dataclass
class InferenceConfig:
    """infer_config.jl:1-16 (MH forces two particles)."""
    num_iters: int
    num_particles: int
    use_dd_proposals: bool = True
    use_lo_sweeps: bool = True
    use_mh_instead_of_pg: bool = False
    rejuv_frequency: int = 50
    reporting_frequency: int = 100

    def __post_init__(self):
        if self.use_mh_instead_of_pg:
            self.num_particles = 2
