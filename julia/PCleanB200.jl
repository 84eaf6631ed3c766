# PCleanB200.jl — ccall shim over libpclean_b200.so (include/pclean_b200.h).
# UNTESTED: no Julia toolchain exists in the build image; the ctypes twin
# (pclean_b200/engine.py) is the tested host.  See INTEGRATION.md.
module PCleanB200

const LIB = get(ENV, "PCLEAN_B200_LIB", "libpclean_b200.so")

struct CValue
    tag::Int32
    i::Int32
    d::Float64
end

struct SweepStats
    rows::Int64; particles::Int64; new_rows::Int64; dummy_draws::Int64; changed_rows::Int64
    sum_log_ml::Float64; kernel_ms::Float32; total_ms::Float32; launches::Int32
end

mutable struct Engine
    h::Ptr{Cvoid}
end

function Engine(num_iters, num_particles; use_mh=false, rejuv_frequency=50, device=0)
    cfg = Int32[num_iters, num_particles, 1, 1, use_mh ? 1 : 0, rejuv_frequency, 100]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:pclean_create, LIB), Int32, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), cfg, device, out)
    rc == 0 || error("pclean_create failed with $rc (no CUDA device? there is no CPU fallback)")
    e = Engine(out[])
    finalizer(x -> ccall((:pclean_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), e)
    return e
end

last_error(e::Engine) = unsafe_string(ccall((:pclean_last_error, LIB), Cstring, (Ptr{Cvoid},), e.h))
check(e::Engine, rc) = rc == 0 ? nothing : error("pclean_b200: $(last_error(e)) ($rc)")

function sweep!(e::Engine, class::Integer, seed::Integer, sweep_idx::Integer)
    st = Ref(SweepStats(0, 0, 0, 0, 0, 0.0, 0f0, 0f0, 0))
    check(e, ccall((:pclean_sweep, LIB), Int32, (Ptr{Cvoid}, Int32, UInt64, UInt32, Ref{SweepStats}),
                   e.h, class, seed, sweep_idx, st))
    return st[]
end

function download_assignment(e::Engine, class, fk_vertex, n_rows)
    keys = Vector{Int64}(undef, n_rows)
    check(e, ccall((:pclean_download_assignment, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int64, Ptr{Int64}),
                   e.h, class, fk_vertex, n_rows, keys))
    return keys
end

end # module
