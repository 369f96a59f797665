# GpuRbf.jl -- the Julia side of the drop-in: a `GpuRbfConfig <: AbstractSurrogateConfig` whose methods reach
# CUDA only through `ccall` into libmorbit_rbf.so (include/morbit_rbf.h).  No CUDA.jl, no kernel DSL.
#
# This file cannot be executed in the build image (no `julia` binary); it is the binding a Morbit maintainer adds
# (`include("GpuRbf.jl")` after `models/RbfModel.jl` in src/Morbit.jl:80).  Every method below names the reference
# method it replaces.  The Python mirror `morbit.jl_b200/surrogate.py` implements exactly the same host logic and is
# what the parity tests drive.
#
# Host logic kept in Julia (as in the reference): database appends (`new_result!`), site matching for
# `_exploit_other_rbf_metas!`, the evaluation budget, string shape parameters.  Everything numerical is one ccall.

const LIBMRBF = get(ENV, "MORBIT_RBF_LIB", "libmorbit_rbf.so")

# struct mrbf_cfg (include/morbit_rbf.h) -- field order and types must match
struct MrbfCfg
    kernel::Int32
    polynomial_degree::Int32
    shape_parameter::Float64
    theta_enlarge_1::Float64
    theta_enlarge_2::Float64
    theta_pivot::Float64
    theta_pivot_cholesky::Float64
    max_model_points::Int32
    use_max_points::Int32
    optimized_sampling::Int32
    reserved::Int32
end

const KERNEL_IDS = Dict(:cubic => 0, :inv_multiquadric => 1, :multiquadric => 2, :thin_plate_spline => 3, :gaussian => 4)

# one context per (Julia thread, device): the reference runs many `optimize` under Threads.@threads
# (examples/large_scale_benchmarks.jl:253), the library is re-entrant per context
const _CTX = Dict{Tuple{Int,Int},Ptr{Cvoid}}()
const _CTX_LOCK = ReentrantLock()
function mrbf_ctx(device::Int = 0)
    lock(_CTX_LOCK) do
        get!(_CTX, (Threads.threadid(), device)) do
            ref = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:mrbf_init, LIBMRBF), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref)
            rc == 0 || error("mrbf_init failed ($rc): no usable CUDA device; there is no CPU fallback")
            ref[]
        end
    end
end
_check(ctx, rc) = rc == 0 || error("libmorbit_rbf: " * unsafe_string(ccall((:mrbf_last_error, LIBMRBF), Cstring, (Ptr{Cvoid},), ctx)))

# ---------------------------------------------------------------------------------------------------------------
# Config / meta / model  (replace RbfConfig, RbfMeta, RbfModel -- src/models/RbfModel.jl:33-38, 66-112, 148-159)
# ---------------------------------------------------------------------------------------------------------------
@with_kw struct GpuRbfConfig <: AbstractSurrogateConfig
    rbf::RbfConfig = RbfConfig()      # all numerical settings are RbfConfig's; asserts run in its constructor
    device::Int = 0
end
max_evals(cfg::GpuRbfConfig) = max_evals(cfg.rbf)                      # AbstractSurrogateInterface.jl:6
combinable(::GpuRbfConfig) = true                                      # :17
Base.hash(cfg::GpuRbfConfig, h::UInt) = hash((cfg.rbf, cfg.device), h) # grouping uses isequal, VecFun.jl:366-371
Base.isequal(a::GpuRbfConfig, b::GpuRbfConfig) = isequal(a.rbf, b.rbf) && a.device == b.device
get_saveable_type(cfg::GpuRbfConfig, x, y) = get_saveable_type(cfg.rbf, x, y)
requires_update(::GpuRbfConfig) = true
requires_improve(::GpuRbfConfig) = true

const GpuRbfMeta = RbfMeta          # same bookkeeping fields (center_index, round1..4_indices, fully_linear, improving_directions)

mutable struct GpuRbfModel <: AbstractSurrogate
    handle::Ptr{Cvoid}              # opaque mrbf_model* (device-resident centres, coefficients)
    ctx::Ptr{Cvoid}
    n::Int
    k::Int
    fully_linear::Bool
    function GpuRbfModel(handle, ctx, n, k, fl)
        m = new(handle, ctx, n, k, fl)
        finalizer(m) do mm
            ccall((:mrbf_free_model, LIBMRBF), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}), mm.ctx, mm.handle)
        end
        m
    end
end
fully_linear(m::GpuRbfModel) = m.fully_linear
set_fully_linear!(m::GpuRbfModel, v) = (m.fully_linear = v; nothing)
num_outputs(m::GpuRbfModel) = m.k

function _c_cfg(cfg::RbfConfig, Δ)
    sp = cfg.shape_parameter isa String ? parse_shape_param_string(Δ, cfg.shape_parameter) : cfg.shape_parameter
    MrbfCfg(KERNEL_IDS[cfg.kernel], cfg.polynomial_degree, Float64(sp), cfg.θ_enlarge_1, cfg.θ_enlarge_2, cfg.θ_pivot,
            cfg.θ_pivot_cholesky, cfg.max_model_points, cfg.use_max_points, cfg.optimized_sampling, 0)
end

# dense column-major n x #db matrix of the sub-database sites (ids are the column numbers)
_site_matrix(db) = reduce(hcat, (Vector{Float64}(get_site(db, id)) for id in get_ids(db)); init = zeros(Float64, length(get_site(db, 1)), 0))

# ---------------------------------------------------------------------------------------------------------------
# prepare_init_model / prepare_update_model   (RbfModel.jl:506-513, 518-655)  -> mrbf_select_points
# ---------------------------------------------------------------------------------------------------------------
function prepare_init_model(cfg::GpuRbfConfig, func_indices, mop, scal, id, sdb, ac; ensure_fully_linear = true, kwargs...)
    F = eltype(get_x_scaled(id))
    meta = RbfMeta{F,typeof(func_indices)}(; signature = _get_signature(cfg.rbf), func_indices)
    prepare_update_model(nothing, meta, cfg, func_indices, mop, scal, id, sdb, ac; ensure_fully_linear, kwargs...)
end

function prepare_update_model(mod::Union{Nothing,GpuRbfModel}, meta::RbfMeta, cfg::GpuRbfConfig, func_indices, mop, scal,
        iter_data, sdb, algo_config; ensure_fully_linear = false, force_rebuild = false, meta_array = nothing)
    rcfg = cfg.rbf
    ctx = mrbf_ctx(cfg.device)
    db = get_sub_db(sdb, func_indices)
    Δ = Float64(get_delta(iter_data)); Δ_max = Float64(delta_max(algo_config))
    x = Vector{Float64}(get_x_scaled(iter_data)); n = length(x)
    x_index = get_x_index(iter_data, Tuple(func_indices))
    meta.fully_linear = false
    skip = _exploit_other_rbf_metas!(meta, db, sdb, meta_array)          # host logic over ids, RbfModel.jl:311-342
    meta.center_index = x_index
    lb, ub = full_bounds_internal(scal)
    ccfg = Ref(_c_cfg(rcfg, Δ))
    sites = _site_matrix(db); n_db = Int32[size(sites, 2)]
    max_points = rcfg.max_model_points <= 0 ? ((n + 1) * (n + 2)) ÷ 2 : rcfg.max_model_points
    r4 = zeros(Int32, max_points); n_r4 = Int32[0]; status = Int32[0]
    if skip                                                               # @goto round4, RbfModel.jl:562
        empty!(meta.round4_indices)
        if rcfg.optimized_sampling
            Δ_2 = rcfg.θ_enlarge_2 * Δ_max
            lb_2 = max.(lb, x .- Δ_2); ub_2 = min.(ub, x .+ Δ_2)
            found = Int32.(_collect_indices(meta)); n_found = Int32[length(found)]
            GC.@preserve sites found lb_2 ub_2 r4 begin
                _check(ctx, ccall((:mrbf_round4, LIBMRBF), Cint,
                    (Ptr{Cvoid}, Ref{MrbfCfg}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                     Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                    ctx, ccfg, 1, n, n_db[1], sites, n_db, lb_2, ub_2, length(found), found, n_found, 0, C_NULL, C_NULL,
                    max_points, r4, n_r4, status))
            end
            append!(meta.round4_indices, Int.(r4[1:n_r4[1]]))
        end
        return meta
    end
    # evaluation budget for round 3, RbfModel.jl:613-618
    num_objf_evals = maximum(num_evals(_get(mop, ind)) for ind in func_indices)
    budget = min(max_evals(algo_config), max_evals(rcfg)) - 1 - num_objf_evals - length(_missing_ids(db))
    max_new = Int32[clamp(budget, 0, typemax(Int32))]
    flags_in = Int32[ensure_fully_linear, force_rebuild]
    r1 = zeros(Int32, n); r2 = zeros(Int32, n); r3 = zeros(Float64, n, n); dirs = zeros(Float64, n, n)
    n_r1 = Int32[0]; n_r2 = Int32[0]; n_r3 = Int32[0]; n_dirs = Int32[0]; flags_out = zeros(Int32, 2)
    xi = Int32[x_index]; Δv = Float64[Δ]
    GC.@preserve sites x lb ub r1 r2 r3 r4 dirs begin
        _check(ctx, ccall((:mrbf_select_points, LIBMRBF), Cint,
            (Ptr{Cvoid}, Ref{MrbfCfg}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64,
             Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
             Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
            ctx, ccfg, 1, n, n_db[1], sites, n_db, xi, x, Δv, Δ_max, Vector{Float64}(lb), Vector{Float64}(ub), flags_in, max_new,
            r1, n_r1, r2, n_r2, r3, n_r3, max_points, r4, n_r4, dirs, n_dirs, flags_out, status))
    end
    F = eltype(get_x_scaled(iter_data))
    empty!(meta.round1_indices); append!(meta.round1_indices, Int.(r1[1:n_r1[1]]))
    empty!(meta.round2_indices); append!(meta.round2_indices, Int.(r2[1:n_r2[1]]))
    empty!(meta.improving_directions); append!(meta.improving_directions, [F.(dirs[:, c]) for c in 1:n_dirs[1]])
    empty!(meta.round3_indices)                      # new_result!(db, p, F[]) -> consecutive ids, RbfModel.jl:301-305
    for i in 1:n_r3[1]
        push!(meta.round3_indices, new_result!(db, F.(r3[:, i]), F[]))
    end
    empty!(meta.round4_indices); append!(meta.round4_indices, Int.(r4[1:n_r4[1]]))
    meta.fully_linear = flags_out[1] != 0
    return meta
end

# prepare_improve_model is n-sized scalar logic on ids and one wall step: the reference method is reused unchanged
prepare_improve_model(mod::Union{Nothing,GpuRbfModel}, meta::RbfMeta, cfg::GpuRbfConfig, args...; kwargs...) =
    prepare_improve_model(nothing, meta, cfg.rbf, args...; kwargs...)      # RbfModel.jl:699-732

# ---------------------------------------------------------------------------------------------------------------
# init_model / update_model / improve_model   (RbfModel.jl:738-776)  -> mrbf_build
# ---------------------------------------------------------------------------------------------------------------
init_model(meta::RbfMeta, cfg::GpuRbfConfig, args...; kwargs...) = update_model(nothing, meta, cfg, args...; kwargs...)
improve_model(mod, meta::RbfMeta, cfg::GpuRbfConfig, args...; kwargs...) = update_model(mod, meta, cfg, args...; kwargs...)

function update_model(mod::Union{Nothing,GpuRbfModel}, meta::RbfMeta, cfg::GpuRbfConfig, func_indices, mop, scal, iter_data, sdb, ac; kwargs...)
    ctx = mrbf_ctx(cfg.device)
    db = get_sub_db(sdb, func_indices)
    Δ = Float64(get_delta(iter_data))
    ids = _collect_indices(meta)                                        # centre, r1, r2, r3, r4 -- RbfModel.jl:178-186
    sites = reduce(hcat, (Vector{Float64}(get_site(db, i)) for i in ids))      # n x N   (column-major == N x n row-major AoS)
    values = reduce(hcat, (Vector{Float64}(get_value(db, i)) for i in ids))    # k x N
    n, N = size(sites); k = size(values, 1)
    handle = Ref{Ptr{Cvoid}}(C_NULL); status = Int32[0]; Nv = Int32[N]
    GC.@preserve sites values begin
        _check(ctx, ccall((:mrbf_build, LIBMRBF), Cint,
            (Ptr{Cvoid}, Ref{MrbfCfg}, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}, Ptr{Int32}),
            ctx, Ref(_c_cfg(cfg.rbf, Δ)), 1, n, k, N, Nv, sites, values, C_NULL, handle, status))
    end
    return GpuRbfModel(handle[], ctx, n, k, meta.fully_linear), meta
end

# ---------------------------------------------------------------------------------------------------------------
# eval_models / get_gradient / get_jacobian   (RbfModel.jl:783-800)  -> mrbf_eval
# ---------------------------------------------------------------------------------------------------------------
function _eval(mod::GpuRbfModel, x̂::Vec; values::Bool, jac::Bool)
    x = Vector{Float64}(x̂)
    Y = values ? zeros(Float64, mod.k) : Float64[]
    J = jac ? zeros(Float64, mod.n, mod.k) : Float64[]          # C layout k x n row-major == Julia n x k column-major
    GC.@preserve x Y J begin
        _check(mod.ctx, ccall((:mrbf_eval, LIBMRBF), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            mod.ctx, mod.handle, 1, x, values ? pointer(Y) : C_NULL, jac ? pointer(J) : C_NULL))
    end
    return Y, J
end
eval_models(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec) = _eval(mod, x̂; values = true, jac = false)[1]
eval_models(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec, ℓ) = eval_models(mod, scal, x̂)[ℓ]       # ℓ::Int or Vector{Int}
get_jacobian(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec, rows = nothing) =
    (J = permutedims(_eval(mod, x̂; values = false, jac = true)[2]); isnothing(rows) ? J : J[rows, :])
get_gradient(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec, ℓ) = vec(get_jacobian(mod, scal, x̂, [ℓ]))

# Armijo backtracking over the surrogate (descent.jl:150-185): one launch for all step sizes instead of <= 118
# sequential model evaluations.  `_backtrack` gains a method for containers whose objectives are one GpuRbfModel.
function _backtrack_gpu(mod::GpuRbfModel, x::Vector{Float64}, dir::Vector{Float64}, step_size, ω, cfg::SteepestDescentConfig)
    n = length(x); k = mod.k
    idx = Int32[0]; σ = Float64[0]; x₊ = zeros(n); mx = zeros(k); mx₊ = zeros(k)
    _check(mod.ctx, ccall((:mrbf_backtrack, LIBMRBF), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Int32, Int32,
         Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        mod.ctx, mod.handle, x, dir, Float64[step_size], Float64[ω], cfg.armijo_const_rhs, cfg.armijo_const_shrink,
        cfg.min_stepsize, cfg.max_loops, cfg.strict_backtracking, idx, σ, x₊, mx, mx₊))
    return x₊, mx₊, σ[1] .* dir
end

# Optional fast path for update_model: `prepare_update_model` may call `mrbf_select_points_keep` (same arguments as
# `mrbf_select_points` plus a `Ref{Ptr{Cvoid}}` for the kept round-4 factorisation, stored in the GpuRbfMeta) and `update_model`
# then calls `mrbf_build_prepared(ctx, cfg, prepared, k, sites, values, r3_sites, r3_values, x_index, r1, n_r1, r2, n_r2, n_r3,
# handle, status)` with the database arrays the selection saw and the freshly evaluated round-3 values: the model is finished with
# two triangular solves instead of a from-scratch factorisation (the reference notes this saving itself, RbfModel.jl:657-660).
# `prepare_improve_model` appends a site and therefore drops the kept factorisation (falls back to `mrbf_build`).

# Constrained steepest-descent direction (descent.jl:91-135): the LP that the reference hands to JuMP + OSQP, solved exactly on
# the device.  Drop-in for `_steepest_descent_direction(x, ∇F, lb, ub, [], [], [], [], normalize)` when the MOP has no linear
# constraints (descent.jl:239 passes them through; with constraints the reference method stays in charge).
function _steepest_descent_direction_gpu(ctx::Ptr{Cvoid}, x::Vector{Float64}, ∇F::Matrix{Float64}, lb::Vector{Float64}, ub::Vector{Float64},
                                         normalize::Bool = true)
    k, n = size(∇F)
    jac = permutedims(∇F)                       # k x n row-major == n x k column-major
    d = zeros(n); ω = Float64[0]; iters = Int32[0]; status = Int32[0]
    _check(ctx, ccall((:mrbf_descent_direction, LIBMRBF), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
         Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
        ctx, 1, n, k, jac, x, lb, ub, normalize, d, ω, iters, status))
    status[1] == 0 || return zeros(n), -Inf     # descent.jl:129-133
    return d, ω[1]
end

# user-facing: add_objective!(mop, f; model_cfg = GpuRbfConfig(rbf = RbfConfig(kernel = :multiquadric)), n_out = 2)
