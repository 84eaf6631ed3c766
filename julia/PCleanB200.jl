# PCleanB200.jl — Julia host shim over the C ABI of pclean_b200 (include/pclean_b200.h).
#
# UNTESTED HERE: the build image and the GPU box have no `julia` binary (SURVEY.md §8c).  The file
# mirrors, function for function, what pclean_b200/lowering.py + engine.py do (those ARE tested, and
# tests/test_irfile.py pins the file format this shim writes), so that a maintainer with Julia can:
#
#     using PClean; include("julia/PCleanB200.jl"); using .PCleanB200
#     trace = PCleanB200.initialize_trace(observations, config)   # replaces PClean.initialize_trace (inference.jl:3)
#     PCleanB200.run_inference!(trace, config)                    # replaces PClean.run_inference!  (inference.jl:83)
#     PCleanB200.materialize!(trace)                              # trace.pclean.tables[c].rows[key][vid], as analysis.jl:44-56 reads them
#     evaluate_accuracy(dirty, clean, trace.pclean.tables[query.class], query)
#
# or, without a GPU at hand, `PCleanB200.write_ir(path, ir, obs)` to hand the model + data to the
# Python / C harness (`Engine(path, cfg)`, `pclean_load_model_file`).
#
# What crosses the boundary: the flat IR of SURVEY App. D (pclean_model_ir), the string dictionary,
# tabulated JuliaNode closures (the device cannot run closures: zero-argument nodes are constants,
# "$(a)<sep>$(b)" and round(unit.backward(x)) are recognised by probing, every other closure is
# tabulated over the finite product of its arguments' supports), the observed dataset, InferenceConfig.
module PCleanB200

using PClean
using PClean: PCleanModel, PCleanClass, PCleanNode, JuliaNode, RandomChoiceNode, ParameterNode, ForeignKeyNode,
              SubmodelNode, ExternalLikelihoodNode, Plan, Query, ObservedDataset, InferenceConfig, PCleanTrace,
              TableTrace, VertexID, ClassID, strip_subnodes, supports_explicitly_missing_observations
using DataFrames: eachrow, propertynames

const LIB = get(ENV, "PCLEAN_B200_LIB", joinpath(@__DIR__, "..", "pclean_b200", "libpclean_b200.so"))

# ---------------------------------------------------------------- values (pclean_b200.h)
const VAL_ABSENT, VAL_MISSING, VAL_STR, VAL_REAL, VAL_INT, VAL_LIST, VAL_XFORM, VAL_PARAM, VAL_IPARAM, VAL_KEY, VAL_DUMMY = Int32.(0:10)
const NODE_JULIA, NODE_CHOICE, NODE_PARAM, NODE_FK = Int32.(0:3)
const WRAP_NONE, WRAP_SUBMODEL, WRAP_EXTERNAL = Int32.(0:2)
const FUNC_CONST, FUNC_TABLE, FUNC_ROUND_BACKWARD, FUNC_JOIN = Int32.(0:3)
const PARAM_PROPORTIONS, PARAM_MEAN, PARAM_PROB = Int32.(0:2)

struct CValue
    tag::Int32
    i::Int32
    d::Float64
end
CValue(tag, i) = CValue(Int32(tag), Int32(i), 0.0)

dist_code(d) = d isa PClean.AddTypos ? 0 : d isa PClean.ChooseProportionally ? 1 : d isa PClean.ChooseUniformly ? 2 :
               d isa PClean.StringPrior ? 3 : d isa PClean.TimePrior ? 4 : d isa PClean.MaybeSwap ? 5 :
               d isa PClean.TransformedGaussian ? 6 : d isa PClean.Unmodeled ? 7 : d isa PClean.AddNoise ? 8 :
               error("distribution $(typeof(d)) is not on the engine's path (DESIGN.md section 8)")

# handles standing for parameter objects while closures are tabulated (lowering.py ParamHandle / ParamSlot)
struct ParamHandle; class::ClassID; vertex::Int; end
struct ParamSlot; class::ClassID; vertex::Int; key::Any; end
Base.getindex(p::ParamHandle, key) = ParamSlot(p.class, p.vertex, key)      # IndexedParameter[key] inside a closure

# ---------------------------------------------------------------- the flat IR
mutable struct FlatIR
    model::PCleanModel
    class_index::Dict{ClassID,Int}
    strings::Vector{String}; string_id::Dict{String,Int}
    lists::Vector{Vector{CValue}}; list_id::Dict{Vector{CValue},Int}
    xforms::Vector{Float64}; xform_id::Dict{Float64,Int}
    param_spec::Dict{Tuple{ClassID,Int},Int}; param_kind::Vector{Int32}; param_indexed::Vector{Int32}
    param_prior::Vector{Tuple{Float64,Float64}}
    slot_param::Vector{Int32}; slot_id::Dict{Tuple{Int,Any},Int}; slot_key::Vector{Any}
    func_id::Dict{UInt,Int}; func_kind::Vector{Int32}; func_const::Vector{CValue}
    func_keyargs::Vector{Vector{Int32}}; func_tab::Vector{Vector{Tuple{Vector{Int32},CValue}}}
    support_cache::Dict{Tuple{ClassID,Int},Union{Nothing,Vector{Any}}}
    data_support::Dict{Tuple{ClassID,Int},Vector{Any}}
    arrays::Dict{String,Any}           # name -> Vector{Int32|Int64|Float64|UInt32|CValue}
    scalars::Dict{String,Int32}
end

function intern!(ir::FlatIR, s::AbstractString)
    get!(ir.string_id, String(s)) do
        push!(ir.strings, String(s)); length(ir.strings) - 1
    end
end

# a Transformation the engine understands is linear: backward(x) = x * scale (transformed_gaussian.jl:5-9)
function xform_scale(t)
    s = t.backward(1.0)
    isapprox(t.backward(2.0), 2.0 * s) && isapprox(t.backward(0.0), 0.0; atol = 1e-12) ||
        error("only linear transformations (backward(x) = x * scale) are on the engine's path")
    Float64(s)
end

function slot!(ir::FlatIR, spec::Int, key)
    get!(ir.slot_id, (spec, key)) do
        push!(ir.slot_param, Int32(spec)); push!(ir.slot_key, key); length(ir.slot_param) - 1
    end
end

function encode!(ir::FlatIR, v)::CValue
    v === missing && return CValue(VAL_MISSING, 0)
    v isa AbstractString && return CValue(VAL_STR, intern!(ir, v))
    v isa Bool && return CValue(VAL_INT, Int(v))
    v isa Integer && return CValue(VAL_INT, v)
    v isa AbstractFloat && return CValue(VAL_REAL, Int32(0), Float64(v))
    if v isa PClean.Transformation
        sc = xform_scale(v)
        k = get!(ir.xform_id, sc) do
            push!(ir.xforms, sc); length(ir.xforms) - 1
        end
        return CValue(VAL_XFORM, k)
    end
    v isa ParamSlot && return CValue(VAL_PARAM, slot!(ir, ir.param_spec[(v.class, v.vertex)], v.key))
    if v isa ParamHandle
        spec = ir.param_spec[(v.class, v.vertex)]
        return ir.param_indexed[spec + 1] != 0 ? CValue(VAL_IPARAM, spec) : CValue(VAL_PARAM, slot!(ir, spec, nothing))
    end
    if v isa AbstractVector || v isa Tuple
        enc = CValue[encode!(ir, x) for x in v]
        k = get!(ir.list_id, enc) do
            push!(ir.lists, enc); length(ir.lists) - 1
        end
        return CValue(VAL_LIST, k)
    end
    error("cannot encode a value of type $(typeof(v)) across the C ABI")
end

# ---------------------------------------------------------------- supports (lowering.py support())
function register!(ir::FlatIR, cls::ClassID, v::Int, vals)
    cur = get!(ir.data_support, (cls, v), Any[])
    for x in vals
        x in cur || push!(cur, x)
    end
    node = ir.model.classes[cls].nodes[v]
    if node isa SubmodelNode
        fk = strip_subnodes(ir.model.classes[cls].nodes[node.foreign_key_node_id])
        register!(ir, fk.target_class, node.subnode_id, vals)
    end
end

function register_data_support!(ir::FlatIR, ds::ObservedDataset)
    q = ds.query
    for (col, v) in vcat(collect(q.obsmap), collect(q.cleanmap))
        col in propertynames(ds.data) || continue
        register!(ir, q.class, v, unique(skipmissing(ds.data[!, col])))
    end
end

flatten_lists(::Nothing) = nothing
flatten_lists(lists) = unique(Iterators.flatten(lists)) |> collect |> x -> Vector{Any}(x)

function support(ir::FlatIR, cls::ClassID, v::Int)
    key = (cls, v)
    haskey(ir.support_cache, key) && return ir.support_cache[key]
    ir.support_cache[key] = nothing                       # cycle guard
    cm = ir.model.classes[cls]
    node = cm.nodes[v]
    out = nothing
    if node isa SubmodelNode
        fk = strip_subnodes(cm.nodes[node.foreign_key_node_id])
        out = support(ir, fk.target_class, node.subnode_id)
    elseif node isa ParameterNode
        out = Any[ParamHandle(cls, v)]
    elseif node isa JuliaNode
        out = julia_outputs(ir, cls, node)
    elseif node isa RandomChoiceNode
        d = node.dist
        if d isa PClean.ChooseProportionally || d isa PClean.ChooseUniformly || d isa PClean.TimePrior
            out = flatten_lists(support(ir, cls, node.arg_node_ids[1]))
        elseif d isa PClean.StringPrior
            out = flatten_lists(support(ir, cls, node.arg_node_ids[3]))
        end
        data = get(ir.data_support, key, nothing)
        if data !== nothing && !isempty(data)
            out = unique(vcat(out === nothing ? Any[] : out, data))
        end
    end
    if out === nothing && haskey(ir.data_support, key)
        out = copy(ir.data_support[key])
    end
    ir.support_cache[key] = out
end

# recognise the two closures the engine evaluates natively (lowering.py `builtin`)
function detect_builtin(f, nargs::Int)
    if nargs == 2
        try
            r = f("q7", "z3")
            if r isa AbstractString && startswith(r, "q7") && endswith(r, "z3")
                sep = r[3:end-2]
                f("ab", "cd") == "ab" * sep * "cd" && return (:join, String(sep))
            end
        catch
        end
        try
            t = PClean.Transformation(x -> x / 3.0, x -> x * 3.0, x -> 1 / 3.0)
            f(t, 7.4) == round(t.backward(7.4)) && f(t, 100.2) == round(t.backward(100.2)) && return (:round_backward, nothing)
        catch
        end
    end
    nothing
end

function julia_outputs(ir::FlatIR, cls::ClassID, node::JuliaNode)
    b = detect_builtin(node.f, length(node.arg_node_ids))
    b !== nothing && b[1] == :round_backward && return nothing
    sups = Vector{Any}[]
    for a in node.arg_node_ids
        s = support(ir, cls, a)
        s === nothing && return nothing
        push!(sups, s)
    end
    prod(max(1, length(s)) for s in sups; init = 1) > 5_000_000 && return nothing
    out = Any[]
    for combo in Iterators.product(sups...)
        r = try node.f(combo...) catch e; (e isa KeyError || e isa BoundsError) ? continue : rethrow() end
        r in out || push!(out, r)
    end
    out
end

key_code(ir::FlatIR, v) = (c = encode!(ir, v); c.tag in (VAL_STR, VAL_INT, VAL_LIST, VAL_XFORM, VAL_PARAM, VAL_IPARAM) ? c.i :
                           error("JuliaNode key argument must be discrete, got $(v)"))

function function!(ir::FlatIR, cls::ClassID, node::JuliaNode)
    id = objectid(node.f)
    haskey(ir.func_id, id) && return ir.func_id[id]
    fid = length(ir.func_kind)
    ir.func_id[id] = fid
    push!(ir.func_kind, FUNC_CONST); push!(ir.func_const, CValue(VAL_ABSENT, 0)); push!(ir.func_keyargs, Int32[]); push!(ir.func_tab, [])
    b = detect_builtin(node.f, length(node.arg_node_ids))
    if b !== nothing
        ir.func_kind[fid + 1] = b[1] == :join ? FUNC_JOIN : FUNC_ROUND_BACKWARD
        b[1] == :join && (ir.func_const[fid + 1] = encode!(ir, b[2]))
        return fid
    end
    if isempty(node.arg_node_ids)
        ir.func_const[fid + 1] = encode!(ir, node.f())
        return fid
    end
    ir.func_kind[fid + 1] = FUNC_TABLE
    sups = Vector{Any}[]; keypos = Int32[]
    for (pos, a) in enumerate(node.arg_node_ids)
        s = support(ir, cls, a)
        s === nothing && error("JuliaNode in class $cls has an argument (vertex $a) with no finite support")
        push!(sups, s)
        (length(s) == 1 && s[1] isa ParamHandle) || push!(keypos, pos - 1)
    end
    ir.func_keyargs[fid + 1] = keypos
    entries = Tuple{Vector{Int32},CValue}[]
    for combo in Iterators.product(sups...)
        r = try node.f(combo...) catch e; (e isa KeyError || e isa BoundsError) ? continue : rethrow() end
        push!(entries, (Int32[key_code(ir, combo[p + 1]) for p in keypos], encode!(ir, r)))
    end
    ir.func_tab[fid + 1] = entries
    fid
end

# tabulate a closure in the class that declared it (SubmodelNode copies share the closure object)
function function_for!(ir::FlatIR, cls::ClassID, node::JuliaNode)
    haskey(ir.func_id, objectid(node.f)) && return ir.func_id[objectid(node.f)]
    for c in ir.model.class_order, n in ir.model.classes[c].nodes
        n isa JuliaNode && n.f === node.f && return function!(ir, c, n)
    end
    function!(ir, cls, node)
end

function flatten_plan!(plan::Plan, pv::Vector{Int32}, pn::Vector{Int32})
    for step in plan.steps
        push!(pv, Int32(step.idx - 1)); push!(pn, Int32(length(step.rest.steps)))
        flatten_plan!(step.rest, pv, pn)
    end
end

"""
    flatten_model(model, datasets) -> FlatIR

App. D flattening of `PCleanModel` (model.jl:87-188); the counterpart of `FlatIR.__init__` in
pclean_b200/lowering.py — same array names, same orders, vertex ids 0-based on the C side.
"""
function flatten_model(model::PCleanModel, datasets::Vector{ObservedDataset})
    ir = FlatIR(model, Dict(c => k - 1 for (k, c) in enumerate(model.class_order)), String[], Dict(), Vector{CValue}[], Dict(),
                Float64[], Dict(), Dict(), Int32[], Int32[], Tuple{Float64,Float64}[], Int32[], Dict(), Any[], Dict(), Int32[], CValue[],
                Vector{Int32}[], [], Dict(), Dict(), Dict(), Dict())
    foreach(ds -> register_data_support!(ir, ds), datasets)
    order = model.class_order
    # parameter specs first (class order, vertex order): slots of basic parameters are stable
    for cls in order, (v, node) in enumerate(model.classes[cls].nodes)
        node isa ParameterNode || continue
        p = node.make_parameter()
        spec = length(ir.param_kind)
        ir.param_spec[(cls, v)] = spec
        indexed = p isa PClean.IndexedParameter
        prior = indexed ? p.shared_prior : p.prior
        basic = indexed ? typeof(p).parameters[2] : typeof(p)
        kind, p0, p1 = basic <: PClean.ProportionsParameter ? (PARAM_PROPORTIONS, prior isa PClean.VariableSizeProportionsParameterPrior ? prior.concentration : prior.concentrations[1], 0.0) :
                       basic <: PClean.MeanParameter ? (PARAM_MEAN, prior.mean, prior.std) :
                       basic <: PClean.ProbParameter ? (PARAM_PROB, prior.a, prior.b) : error("parameter type $basic")
        push!(ir.param_kind, kind); push!(ir.param_indexed, Int32(indexed)); push!(ir.param_prior, (Float64(p0), Float64(p1)))
        indexed || slot!(ir, spec, nothing)
    end
    A = Dict{String,Vector{Int32}}(n => Int32[] for n in ("v_kind", "v_wrap", "wrap_fk", "wrap_subid", "v_dist", "v_args", "v_func", "v_target",
        "v_vmap", "v_param", "v_path", "v_extv", "block_v", "plan_vertex", "plan_nchild", "hash_v", "path_target", "path_class", "path_vertex", "path_vmap"))
    for n in ("class_voff", "v_wrap_off", "v_args_off", "v_vmap_off", "class_block_off", "block_voff", "plan_off", "class_hash_off", "path_len_off", "path_vmap_off")
        A[n] = Int32[0]
    end
    path_ids = Dict{Tuple{ClassID,Any},Int}()
    for cls in order
        cm = model.classes[cls]
        n_normal = count(n -> !(n isa ExternalLikelihoodNode), cm.nodes)
        for (path, vmap) in cm.incoming_references
            path_ids[(cls, path)] = length(A["path_target"])
            push!(A["path_target"], ir.class_index[cls])
            for av in path
                push!(A["path_class"], ir.class_index[av.class]); push!(A["path_vertex"], av.node_id - 1)
            end
            push!(A["path_len_off"], length(A["path_class"]))
            dense = fill(Int32(-1), n_normal)
            for (i, j) in vmap; dense[i] = j - 1; end
            append!(A["path_vmap"], dense); push!(A["path_vmap_off"], length(A["path_vmap"]))
        end
    end
    for cls in order
        cm = model.classes[cls]
        for (v, node) in enumerate(cm.nodes)
            wrap, base, ext_path, ext_v, func_cls = WRAP_NONE, node, -1, -1, cls
            if node isa ExternalLikelihoodNode
                wrap, base = WRAP_EXTERNAL, node.external_node
                ext_path, ext_v, func_cls = path_ids[(cls, node.path)], node.external_node_id - 1, node.path[end].class
            elseif node isa SubmodelNode
                wrap = WRAP_SUBMODEL
                while base isa SubmodelNode
                    push!(A["wrap_fk"], base.foreign_key_node_id - 1); push!(A["wrap_subid"], base.subnode_id - 1)
                    base = base.subnode
                end
            end
            push!(A["v_wrap_off"], length(A["wrap_fk"])); push!(A["v_wrap"], wrap); push!(A["v_path"], ext_path); push!(A["v_extv"], ext_v)
            kind, dist, func, target, param = -1, -1, -1, -1, -1
            if base isa JuliaNode
                kind = NODE_JULIA; append!(A["v_args"], base.arg_node_ids .- 1); func = function_for!(ir, func_cls, base)
            elseif base isa RandomChoiceNode
                kind = NODE_CHOICE; dist = dist_code(base.dist); append!(A["v_args"], base.arg_node_ids .- 1)
            elseif base isa ParameterNode
                kind = NODE_PARAM
                param = first(s for ((c, vv), s) in ir.param_spec if model.classes[c].nodes[vv] === base)
            elseif base isa ForeignKeyNode
                kind = NODE_FK; target = ir.class_index[base.target_class]
                append!(A["v_vmap"], Int32[base.vmap[i] - 1 for i in 1:length(base.vmap)])
            end
            push!(A["v_kind"], kind); push!(A["v_dist"], dist); push!(A["v_func"], func); push!(A["v_target"], target); push!(A["v_param"], param)
            push!(A["v_args_off"], length(A["v_args"])); push!(A["v_vmap_off"], length(A["v_vmap"]))
        end
        push!(A["class_voff"], length(A["v_kind"]))
        for (block, plan) in zip(cm.blocks, cm.plans)
            append!(A["block_v"], block .- 1); push!(A["block_voff"], length(A["block_v"]))
            flatten_plan!(plan, A["plan_vertex"], A["plan_nchild"]); push!(A["plan_off"], length(A["plan_vertex"]))
        end
        push!(A["class_block_off"], length(A["block_voff"]) - 1)
        append!(A["hash_v"], cm.hash_keys .- 1); push!(A["class_hash_off"], length(A["hash_v"]))
    end
    for (k, v) in A; ir.arrays[k] = v; end
    ir.arrays["py_strength"] = Float64[model.classes[c].initial_pitman_yor_params.strength for c in order]
    ir.arrays["py_discount"] = Float64[model.classes[c].initial_pitman_yor_params.discount for c in order]
    ir.scalars["n_classes"] = length(order); ir.scalars["n_vertices"] = length(A["v_kind"])
    ir.scalars["n_blocks"] = length(A["block_voff"]) - 1; ir.scalars["n_paths"] = length(A["path_target"])
    finalize_tables!(ir)
    ir
end

nonempty(v::Vector{T}, z::T) where {T} = isempty(v) ? T[z] : v

# dictionary / list / function arrays; called again after new values were interned (lowering.py refresh())
function finalize_tables!(ir::FlatIR)
    a = ir.arrays
    a["func_kind"] = copy(ir.func_kind); a["func_const"] = nonempty(copy(ir.func_const), CValue(VAL_ABSENT, 0))
    ko, ka, to, tk, tko, tv = Int32[0], Int32[], Int32[0], Int32[], Int64[0], CValue[]
    for fid in 1:length(ir.func_kind)
        append!(ka, ir.func_keyargs[fid]); push!(ko, length(ka))
        for (key, val) in ir.func_tab[fid]
            append!(tk, key); push!(tko, length(tk)); push!(tv, val)
        end
        push!(to, length(tv))
    end
    a["func_keyarg_off"] = ko; a["func_keyargs"] = nonempty(ka, Int32(0)); a["func_tab_off"] = to
    a["tab_keys"] = nonempty(tk, Int32(0)); a["tab_key_off"] = tko; a["tab_vals"] = nonempty(tv, CValue(VAL_ABSENT, 0))
    a["param_kind"] = nonempty(copy(ir.param_kind), Int32(0)); a["param_indexed"] = nonempty(copy(ir.param_indexed), Int32(0))
    a["param_prior0"] = nonempty(Float64[p[1] for p in ir.param_prior], 0.0); a["param_prior1"] = nonempty(Float64[p[2] for p in ir.param_prior], 0.0)
    a["slot_param"] = nonempty(copy(ir.slot_param), Int32(0))
    lo, lv = Int64[0], CValue[]
    for l in ir.lists; append!(lv, l); push!(lo, length(lv)); end
    a["list_off"] = lo; a["list_vals"] = nonempty(lv, CValue(VAL_ABSENT, 0))
    a["xform_scale"] = nonempty(copy(ir.xforms), 1.0)
    so, cp = Int64[0], UInt32[]
    for s in ir.strings; append!(cp, UInt32.(collect(s))); push!(so, length(cp)); end
    a["str_off"] = so; a["str_cp"] = nonempty(cp, UInt32(0))
    a["lm_unigram"] = Float64.(vec(PClean.initial_letter_probs))
    a["lm_bigram"] = Float64.(vec(permutedims(PClean.english_letter_transitions)))      # row-major [next][prev] (string_prior.jl:8-9)
    ir.scalars["n_funcs"] = length(ir.func_kind); ir.scalars["n_params"] = length(ir.param_kind); ir.scalars["n_param_slots"] = length(ir.slot_param)
    ir.scalars["n_lists"] = length(ir.lists); ir.scalars["n_xforms"] = length(ir.xforms); ir.scalars["n_strings"] = length(ir.strings)
    ir
end

"""
    encode_observations(ir, ds) -> (cls, n_rows, vertex_of_col, cells, columns)

inference.jl:17-33: per-row observation cells, column-major; ABSENT = not an observation of the row,
MISSING = an explicit `missing` the distribution models (add_typos.jl:7, maybe_swap.jl:3).
"""
function encode_observations(ir::FlatIR, ds::ObservedDataset)
    q = ds.query
    cm = ir.model.classes[q.class]
    cols = [c for c in propertynames(ds.data) if haskey(q.obsmap, c)]
    n = length(eachrow(ds.data))
    cells = Vector{CValue}(undef, length(cols) * n)
    voc = Int32[]
    for (ci, col) in enumerate(cols)
        node_id = q.obsmap[col]
        push!(voc, node_id - 1)
        base = strip_subnodes(cm.nodes[node_id])
        explicit = node_id != q.cleanmap[col] && base isa RandomChoiceNode && supports_explicitly_missing_observations(base.dist)
        for (r, x) in enumerate(ds.data[!, col])
            cells[(ci - 1) * n + r] = ismissing(x) ? CValue(explicit ? VAL_MISSING : VAL_ABSENT, 0) : encode!(ir, x)
        end
    end
    finalize_tables!(ir)
    (Int32(ir.class_index[q.class]), Int64(n), voc, cells, String.(cols))
end

# ---------------------------------------------------------------- PCLIRv1 files (pclean_b200/irfile.py)
const SCALARS = ("n_classes", "n_vertices", "n_blocks", "n_paths", "n_funcs", "n_params", "n_param_slots", "n_lists", "n_xforms", "n_strings")
dtype_code(::Vector{Int32}) = 0; dtype_code(::Vector{Int64}) = 1; dtype_code(::Vector{Float64}) = 2
dtype_code(::Vector{UInt32}) = 3; dtype_code(::Vector{CValue}) = 4; dtype_code(::Vector{UInt8}) = 5

function write_ir(path::AbstractString, ir::FlatIR, obs = nothing)
    entries = Pair{String,Any}[]
    for s in SCALARS; push!(entries, s => Int32[ir.scalars[s]]); end
    for (k, v) in ir.arrays; push!(entries, k => v); end
    push!(entries, "class_names" => Vector{UInt8}(join(String.(ir.model.class_order), "\n")))
    if obs !== nothing
        cls, n, voc, cells, cols = obs
        append!(entries, ["obs.cls" => Int32[cls], "obs.n_rows" => Int64[n], "obs.n_cols" => Int32[length(voc)], "obs.vertex_of_col" => voc,
                          "obs.cells" => cells, "obs.columns" => Vector{UInt8}(join(cols, "\n"))])
    end
    open(path, "w") do f
        write(f, "PCLIRv1\n"); write(f, UInt32(length(entries)))
        for (name, arr) in entries
            nb = Vector{UInt8}(name)
            write(f, UInt32(length(nb))); write(f, nb); write(f, UInt32(dtype_code(arr))); write(f, UInt64(length(arr)))
            write(f, zeros(UInt8, mod(-position(f), 8)))
            write(f, arr)                       # CValue is an isbits struct with C layout (16 bytes)
        end
    end
end

# ---------------------------------------------------------------- the C ABI
struct CConfig
    num_iters::Int32; num_particles::Int32; use_dd_proposals::Int32; use_lo_sweeps::Int32
    use_mh_instead_of_pg::Int32; rejuv_frequency::Int32; reporting_frequency::Int32
end
CConfig(c::InferenceConfig) = CConfig(c.num_iters, c.num_particles, c.use_dd_proposals, c.use_lo_sweeps, c.use_mh_instead_of_pg, c.rejuv_frequency, c.reporting_frequency)

mutable struct SweepStats
    rows::Int64; particles::Int64; new_rows::Int64; dummy_draws::Int64; changed_rows::Int64
    sum_log_ml::Float64; kernel_ms::Float32; total_ms::Float32; launches::Int32
    SweepStats() = new(0, 0, 0, 0, 0, 0.0, 0f0, 0f0, 0)
end

mutable struct B200Trace
    handle::Ptr{Cvoid}
    ir::FlatIR
    observations::Vector{ObservedDataset}
    pclean::PCleanTrace                 # filled by materialize!
    seed::UInt64
end

function check(h, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:pclean_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("pclean_b200 error $rc: $msg")
end

"""
    initialize_trace(observations, config; device = 0, seed = rand(UInt64)) -> B200Trace

Drop-in for `PClean.initialize_trace` (inference.jl:3-58): flattens the model, uploads it with the
dataset and runs the SMC initialisation on the device.  One observed dataset (the engine's limit).
"""
function initialize_trace(observations::Vector{ObservedDataset}, config::InferenceConfig; device::Integer = 0, seed::UInt64 = rand(UInt64))
    length(observations) == 1 || error("the engine takes one observed dataset")
    model = first(observations).query.model
    ir = flatten_model(model, observations)
    cls, n, voc, cells, _ = encode_observations(ir, observations[1])
    path = tempname() * ".pclir"
    write_ir(path, ir, (cls, n, voc, cells, String[]))          # the file is the one tested exchange format
    h = Ref{Ptr{Cvoid}}(C_NULL)
    cfg = Ref(CConfig(config))
    rc = ccall((:pclean_create, LIB), Int32, (Ref{CConfig}, Int32, Ref{Ptr{Cvoid}}), cfg, Int32(device), h)
    rc == 0 || error("pclean_create failed ($rc): is a CUDA device visible? (there is no CPU fallback)")
    check(h[], ccall((:pclean_load_model_file, LIB), Int32, (Ptr{Cvoid}, Cstring), h[], path))
    check(h[], ccall((:pclean_load_observations_file, LIB), Int32, (Ptr{Cvoid}, Cstring), h[], path))
    rm(path; force = true)
    check(h[], ccall((:pclean_init_trace, LIB), Int32, (Ptr{Cvoid}, UInt64), h[], seed))
    t = B200Trace(h[], ir, observations, PCleanTrace(model, Dict()), seed)
    finalizer(x -> ccall((:pclean_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), t)
    t
end

"`PClean.run_inference!` (inference.jl:83-88): `config.num_iters` sweeps over every class, on the device"
function run_inference!(trace::B200Trace, config::InferenceConfig)
    st = SweepStats()
    check(trace.handle, ccall((:pclean_run_inference, LIB), Int32, (Ptr{Cvoid}, UInt64, Ref{SweepStats}), trace.handle, trace.seed, st))
    st
end

"""
    update_observations!(trace, sid_cols, real_cols; rows = 1:n)

Send observed cells again from the host's encoded columns (`pclean_update_observations`,
include/pclean_b200.h): `sid_cols[c]` a `Vector{Int32}` of dictionary ids (`-1` = missing) or `nothing`,
`real_cols[c]` a `Vector{Float64}` or `nothing`, one entry per dataset column in the order of
`encode_observations`.  Returns the bytes copied host -> device.
"""
function update_observations!(trace::B200Trace, sid_cols::Vector, real_cols::Vector; rows::UnitRange{Int} = 1:0)
    nc = length(sid_cols)
    sp = Ptr{Int32}[c === nothing ? Ptr{Int32}(C_NULL) : pointer(c) for c in sid_cols]
    rp = Ptr{Float64}[c === nothing ? Ptr{Float64}(C_NULL) : pointer(c) for c in real_cols]
    n = maximum(length(c) for c in vcat(sid_cols, real_cols) if c !== nothing)
    r0, r1 = isempty(rows) ? (0, n) : (first(rows) - 1, last(rows))
    bytes = Ref{Int64}(0)
    GC.@preserve sid_cols real_cols begin
        check(trace.handle, ccall((:pclean_update_observations, LIB), Int32,
                                  (Ptr{Cvoid}, Int32, Ptr{Ptr{Int32}}, Ptr{Ptr{Float64}}, Int64, Int64, Ref{Int64}),
                                  trace.handle, Int32(nc), sp, rp, Int64(r0), Int64(r1), bytes))
    end
    return bytes[]
end

function engine_string(h, sid)
    n = Ref{Int32}(0)
    ccall((:pclean_get_string, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt32}, Ref{Int32}), h, sid, 0, C_NULL, n)
    buf = Vector{UInt32}(undef, max(1, n[]))
    ccall((:pclean_get_string, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt32}, Ref{Int32}), h, sid, n[], buf, n)
    String(Char.(buf[1:n[]]))
end

"""
    materialize!(trace) -> PCleanTrace

Write the device trace back into `trace.pclean.tables[class].rows[i][vertex_id]` for every vertex the
query cleans (what `evaluate_accuracy` / `save_results` read, analysis.jl:15-88).
"""
function materialize!(trace::B200Trace)
    ds = trace.observations[1]; q = ds.query
    n = length(eachrow(ds.data))
    verts = sort(unique(collect(values(q.cleanmap))))
    out = Vector{CValue}(undef, length(verts) * n)
    check(trace.handle, ccall((:pclean_download_cells, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int64, Ptr{CValue}),
                              trace.handle, trace.ir.class_index[q.class], length(verts), Int32.(verts .- 1), n, out))
    cm = q.model.classes[q.class]
    tt = TableTrace(cm.initial_pitman_yor_params, Dict(), Dict(), Dict(), Dict(), Dict(), Dict(), Dict(), Ref(0))
    cache = Dict{Int32,String}()
    for r in 1:n
        row = Dict{VertexID,Any}()
        for (k, v) in enumerate(verts)
            c = out[(k - 1) * n + r]
            c.tag == VAL_STR && (row[v] = get!(() -> engine_string(trace.handle, c.i), cache, c.i))
            c.tag == VAL_REAL && (row[v] = c.d)
        end
        tt.rows[r] = row
    end
    trace.pclean.tables[q.class] = tt
    trace.pclean
end

end # module
