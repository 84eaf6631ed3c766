# Dump reference-held fixtures that pin the CPU oracle (oracle/pclean_oracle.cpp) to the reference.
# usage: julia --project=<PClean> dump_fixtures.jl <PClean checkout> <out dir>
#
# The reference draws from Julia's unseeded global RNG, so sampled values cannot be compared; what CAN
# is everything that is a deterministic function of (model, data, trace):
#   1. densities.json   log-densities of every distribution on the path, on pairs taken from the real data
#                       (AddTypos incl. the third-party edit distance, StringPrior, TimePrior, MaybeSwap,
#                       TransformedGaussian, ChooseProportionally with a fixed parameter vector)
#   2. <exp>_trace.json the trace after initialize_trace (tables: key -> row cells, reference counts,
#                       Pitman-Yor parameters, parameter values) — the oracle installs it as a snapshot —
#   3. <exp>_marginals.json for every observation row: the log-marginal each compiled block proposal
#                       returns for that row given that trace (the first return value of the generated
#                       function, proposal_compiler.jl:405-414: exact enumeration, no randomness in it)
#                       and the CRP prior vector of the block's reference table (trace.jl:53-61).
# tests/test_reference_fixtures.py compares the oracle's enumeration with (3) on the trace of (2).
using PClean, DataFrames, CSV
using PClean: logdensity, AddTypos, StringPrior, TimePrior, MaybeSwap, TransformedGaussian, InferenceConfig,
              initialize_trace, ProposalRowState, compile_proposal, pitman_yor_prior_logprobs
ref, out = ARGS[1], ARGS[2]
mkpath(out)

jstr(x::AbstractString) = "\"" * escape_string(x) * "\""
jstr(x::Symbol) = jstr(String(x))
jstr(x::Missing) = "null"
jstr(x::Nothing) = "null"
jstr(x::Bool) = x ? "true" : "false"
jstr(x::Real) = isfinite(x) ? repr(Float64(x)) : (x > 0 ? "1e999" : "-1e999")
jstr(x::Integer) = string(x)
jstr(x::AbstractVector) = "[" * join(jstr.(x), ", ") * "]"
jstr(x::Tuple) = jstr(collect(x))
jstr(x::AbstractDict) = "{" * join((jstr(string(k)) * ": " * jstr(v) for (k, v) in x), ", ") * "}"
jstr(x) = jstr(string(x))

# ---- 1. densities on real string pairs
hosp = CSV.File(joinpath(ref, "datasets", "hospital_dirty.csv")) |> DataFrame
hclean = CSV.File(joinpath(ref, "datasets", "hospital_clean.csv")) |> DataFrame
dens = Dict{String,Any}("addtypos" => [], "stringprior" => [], "timeprior" => [], "maybeswap" => [], "transformed_gaussian" => [])
for col in (:HospitalName, :Address1, :City, :CountyName, :PhoneNumber, :MeasureName), r in 1:20:nrow(hosp)
    o, w = string(hosp[r, col]), string(hclean[r, col])
    push!(dens["addtypos"], Dict("observed" => o, "word" => w, "logdensity" => logdensity(AddTypos(), o, w),
                                 "other_word" => string(hclean[mod1(r + 7, nrow(hclean)), col]),
                                 "logdensity_other" => logdensity(AddTypos(), o, string(hclean[mod1(r + 7, nrow(hclean)), col])),
                                 "max2" => logdensity(AddTypos(), o, w, 2)))
    push!(dens["stringprior"], Dict("s" => w, "min" => 3, "max" => 40, "logdensity" => logdensity(StringPrior(), w, 3, 40, String[])))
end
for t in ("7:10 a.m.", "12:05 p.m.", "not a time"), opts in (["7:10 a.m."], ["7:10 a.m.", "7:21 a.m.", "bad"])
    o, lp = PClean.discrete_proposal(TimePrior(), opts)
    push!(dens["timeprior"], Dict("options" => opts, "proposal_logprobs" => lp, "logdensity" => logdensity(TimePrior(), t, opts)))
end
for (obs, val, opts, p) in (("a", "a", ["a", "b", "c"], 0.1), ("b", "a", ["a", "b", "c"], 0.1), (missing, "a", ["a", "b"], 0.3), (missing, "z", ["a", "b"], 0.3))
    push!(dens["maybeswap"], Dict("observed" => obs, "val" => val, "options" => opts, "prob" => p, "logdensity" => logdensity(MaybeSwap(), obs, val, opts, p)))
end
open(joinpath(out, "densities.json"), "w") do f; println(f, jstr(dens)); end

# ---- 2 + 3. traces and per-row block marginals
function dump_experiment(name)
    m = Module(Symbol("fx_", name))
    src = read(joinpath(ref, "experiments", name, "run.jl"), String)
    # stop the driver right after initialize_trace: keep `trace`, `query`, `observations`, `config`
    cut = findfirst(r"run_inference!", src)
    src = cut === nothing ? src : src[1:first(cut)-1]
    Core.eval(m, :(using PClean; using DataFrames; using CSV; include(x) = Base.include($m, joinpath($ref, "experiments", $name, x))))
    cd(joinpath(ref, "experiments", name)) do
        Base.include_string(m, src, "run.jl")
    end
    trace = Core.eval(m, :trace); query = Core.eval(m, :query)
    model = trace.model
    tables = Dict{String,Any}()
    for (cls, t) in trace.tables
        tables[String(cls)] = Dict("rows" => Dict(string(k) => Dict(string(v) => x for (v, x) in row if !(x isa PClean.Parameter)) for (k, row) in t.rows),
                                   "reference_counts" => Dict(string(k) => c for (k, c) in t.reference_counts),
                                   "strength" => t.pitman_yor_params.strength, "discount" => t.pitman_yor_params.discount,
                                   "parameters" => Dict(string(v) => (p isa PClean.IndexedParameter ? Dict(string(i) => q.current_value for (i, q) in p.parameters) : p.current_value)
                                                        for (v, p) in t.parameters))
    end
    open(joinpath(out, "$(name)_trace.json"), "w") do f; println(f, jstr(Dict("class_order" => String.(model.class_order), "tables" => tables))); end
    # log-marginal of every block of every observation row, against the trace with that row removed
    cls = query.class; cm = model.classes[cls]; t = trace.tables[cls]
    marg = Dict{String,Any}()
    for key in sort(collect(keys(t.rows)))[1:min(end, 400)]
        row = t.rows[key]
        PClean.unincorporate_row!(trace, cls, key)
        state = ProposalRowState(trace, cls, PClean.initialize_row_trace_for_smc(trace, cls, key), Dict(), nothing)
        per_block = Float64[]
        for b in 1:length(cm.blocks)
            present = Set(v for v in keys(state.row_trace))
            f = get!(() -> compile_proposal(cm, b, present), cm.compiled_proposals[b], present)
            logmarginal, proposed, _ = Base.invokelatest(f, state)
            push!(per_block, logmarginal)
            merge!(state.row_trace, proposed)               # later blocks see what this one sampled (the marginal of block b+1 depends on it)
        end
        marg[string(key)] = per_block
        t.rows[key] = row
        PClean.incorporate_row!(trace, cls, key)
    end
    open(joinpath(out, "$(name)_marginals.json"), "w") do f; println(f, jstr(marg)); end
end
foreach(dump_experiment, ("hospital", "flights"))
