# Time the reference's own experiment drivers (experiments/*/run.jl) section by section.
# usage: julia --project=<PClean> time_experiments.jl <PClean checkout> <out.json>
using PClean, Dates
ref, out = ARGS[1], ARGS[2]
results = Dict{String,Any}("julia" => string(VERSION), "threads" => Threads.nthreads(), "when" => string(now()),
                           "cpu" => Sys.cpu_info()[1].model, "logical_cores" => Sys.CPU_THREADS)
for name in ("hospital", "rents", "flights")
    # each run.jl loads its data, builds the model, then calls initialize_trace / run_inference! / evaluate_accuracy
    src = read(joinpath(ref, "experiments", name, "run.jl"), String)
    # wrap the two calls on the path with @elapsed (textual, so the drivers stay unmodified on disk)
    src = replace(src, r"(\w+)\s*=\s*initialize_trace\(([^\n]*)\)" => s"\1 = begin global __t_init = @elapsed(__tr = initialize_trace(\2)); __tr end")
    src = replace(src, r"run_inference!\(([^\n]*)\)" => s"global __t_sweeps = @elapsed(run_inference!(\1))")
    m = Module(Symbol("ref_", name))
    Core.eval(m, :(using PClean; using DataFrames; using CSV; include(x) = Base.include($m, joinpath($ref, "experiments", $name, x))))
    cd(joinpath(ref, "experiments", name)) do
        t_total = @elapsed Base.include_string(m, src, "run.jl")
        results[name] = Dict("init_s" => Core.eval(m, :(isdefined($m, :__t_init) ? __t_init : nothing)),
                             "sweeps_s" => Core.eval(m, :(isdefined($m, :__t_sweeps) ? __t_sweeps : nothing)),
                             "total_s_including_jit" => t_total)
    end
end
open(out, "w") do f
    print(f, "{")
    first = true
    for (k, v) in results
        first || print(f, ", "); first = false
        print(f, repr(k), ": ", v isa Dict ? "{" * join(("$(repr(a)): $(b === nothing ? "null" : b)" for (a, b) in v), ", ") * "}" : repr(v))
    end
    println(f, "}")
end
