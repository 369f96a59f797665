# GpuRbf.jl -- the Julia side of the drop-in: a `GpuRbfConfig <: AbstractSurrogateConfig` whose methods reach
# CUDA only through `ccall` into libmorbit_rbf.so (include/morbit_rbf.h).  No CUDA.jl, no kernel DSL, no CPU fallback.
#
# This file cannot be executed in the build image (no `julia` binary); it is the binding a Morbit maintainer adds
# (`include("GpuRbf.jl")` after `models/RbfModel.jl` in src/Morbit.jl:80).  Every method below names the reference
# method it replaces.  The Python mirror `morbit.jl_b200/surrogate.py` implements exactly the same host logic and is
# what the parity tests drive; `tests/test_cpu_host.py::test_julia_shim_ccall_signatures_match_header` checks every
# `ccall` argument tuple in this file against the C header's prototypes.
#
# Host logic kept in Julia (as in the reference): database appends (`new_result!`), site matching for
# `_exploit_other_rbf_metas!`, the evaluation budget, string shape parameters.  Everything numerical is one ccall.

const LIBMRBF = get(ENV, "MORBIT_RBF_LIB", "libmorbit_rbf.so")

const MRBF_OK = Cint(0)
const MRBF_ENUMERIC = Cint(-5)      # some instances failed numerically: the handle is valid, status[] says which

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

# ---------------------------------------------------------------------------------------------------------------
# Contexts.  An mrbf_ctx is NOT thread safe (grow-only workspaces, one stream, one error string), and Julia tasks
# migrate between threads (`Threads.@threads` is :dynamic; examples/large_scale_benchmarks.jl:253 runs many `optimize`
# concurrently), so a context can be keyed neither by thread id nor by task.  Instead every device has a POOL of
# contexts: a call borrows one (creating it when the pool is empty), runs, and puts it back.  Concurrent calls get
# different contexts and therefore different streams; a model handle may be used with any context of its device
# (the host-pointer entry points return only after their stream has drained).
# ---------------------------------------------------------------------------------------------------------------
const _CTX_POOL = Dict{Int,Vector{Ptr{Cvoid}}}()
const _CTX_LOCK = ReentrantLock()

function _new_ctx(device::Int)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mrbf_init, LIBMRBF), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref)
    rc == MRBF_OK || error("mrbf_init failed ($rc): no usable CUDA device; there is no CPU fallback")
    return ref[]
end

"Borrow a context of `device` for the duration of `f(ctx)`."
function with_ctx(f, device::Int = 0)
    ctx = lock(_CTX_LOCK) do
        pool = get!(_CTX_POOL, device) do; Ptr{Cvoid}[]; end
        isempty(pool) ? C_NULL : pop!(pool)
    end
    ctx == C_NULL && (ctx = _new_ctx(device))
    try
        return f(ctx)
    finally
        lock(_CTX_LOCK) do; push!(_CTX_POOL[device], ctx); end
    end
end

_errmsg(ctx) = unsafe_string(ccall((:mrbf_last_error, LIBMRBF), Cstring, (Ptr{Cvoid},), ctx))
_check(ctx, rc) = rc == MRBF_OK || error("libmorbit_rbf: " * _errmsg(ctx))

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

const GpuRbfMeta = RbfMeta          # same bookkeeping fields, so `other_meta isa RbfMeta` in _exploit_other_rbf_metas! holds

mutable struct GpuRbfModel <: AbstractSurrogate
    handle::Ptr{Cvoid}              # opaque mrbf_model* (device-resident centres, coefficients)
    device::Int
    n::Int
    k::Int
    fully_linear::Bool
    status::Int32                   # 0, or > 0: the reduced kernel matrix was not positive definite (duplicated sites)
    function GpuRbfModel(handle, device, n, k, fl, status = Int32(0))
        m = new(handle, device, n, k, fl, status)
        # finalizers must not take locks or switch tasks: mrbf_free_model accepts a NULL context (plain cudaFree)
        finalizer(mm -> (mm.handle == C_NULL || ccall((:mrbf_free_model, LIBMRBF), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}), C_NULL, mm.handle); mm.handle = C_NULL), m)
        m
    end
end
fully_linear(m::GpuRbfModel) = m.fully_linear
set_fully_linear!(m::GpuRbfModel, v) = (m.fully_linear = v; nothing)
num_outputs(m::GpuRbfModel) = m.k

# What the reference throws away between prepare_update_model and update_model and notes it should keep
# (RbfModel.jl:657-660): the round-4 factorisation (opaque mrbf_prepared*, device resident) and the selection outputs
# it belongs to.  RbfMeta has no spare field, so the state hangs off the meta object in a weak-keyed side table.
mutable struct KeptSelection
    prepared::Ptr{Cvoid}
    n_db::Int                       # database size the selection saw (ids 1..n_db)
    x_index::Int32
    r1::Vector{Int32}; n_r1::Vector{Int32}
    r2::Vector{Int32}; n_r2::Vector{Int32}
    r3::Matrix{Float64}; n_r3::Vector{Int32}
    function KeptSelection(args...)
        k = new(args...)
        finalizer(kk -> (kk.prepared == C_NULL || ccall((:mrbf_free_prepared, LIBMRBF), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}), C_NULL, kk.prepared); kk.prepared = C_NULL), k)
        k
    end
end
const _KEPT = WeakKeyDict{Any,KeptSelection}()
const _KEPT_LOCK = ReentrantLock()
_kept(meta) = lock(() -> get(_KEPT, meta, nothing), _KEPT_LOCK)
_keep!(meta, k) = lock(() -> (_KEPT[meta] = k), _KEPT_LOCK)
_drop_kept!(meta) = lock(() -> delete!(_KEPT, meta), _KEPT_LOCK)

function _c_cfg(cfg::RbfConfig, Δ)
    sp = cfg.shape_parameter isa String ? parse_shape_param_string(Δ, cfg.shape_parameter) : cfg.shape_parameter
    MrbfCfg(KERNEL_IDS[cfg.kernel], cfg.polynomial_degree, Float64(sp), cfg.θ_enlarge_1, cfg.θ_enlarge_2, cfg.θ_pivot,
            cfg.θ_pivot_cholesky, cfg.max_model_points, cfg.use_max_points, cfg.optimized_sampling, 0)
end

# dense column-major n x #db matrix of the sub-database sites (ids are the column numbers)
_site_matrix(db) = reduce(hcat, (Vector{Float64}(get_site(db, id)) for id in get_ids(db)); init = zeros(Float64, length(get_site(db, 1)), 0))

# ---------------------------------------------------------------------------------------------------------------
# prepare_init_model / prepare_update_model   (RbfModel.jl:506-513, 518-655)  -> mrbf_select_points_keep
# ---------------------------------------------------------------------------------------------------------------
function prepare_init_model(cfg::GpuRbfConfig, func_indices, mop, scal, id, sdb, ac; ensure_fully_linear = true, kwargs...)
    F = eltype(get_x_scaled(id))
    meta = RbfMeta{F,typeof(func_indices)}(; signature = _get_signature(cfg.rbf), func_indices)
    prepare_update_model(nothing, meta, cfg, func_indices, mop, scal, id, sdb, ac; ensure_fully_linear, kwargs...)
end

function prepare_update_model(mod::Union{Nothing,GpuRbfModel}, meta::RbfMeta, cfg::GpuRbfConfig, func_indices, mop, scal,
        iter_data, sdb, algo_config; ensure_fully_linear = false, force_rebuild = false, meta_array = nothing)
    rcfg = cfg.rbf
    db = get_sub_db(sdb, func_indices)
    Δ = Float64(get_delta(iter_data)); Δ_max = Float64(delta_max(algo_config))
    # `Δ ≈ Δ_max` at RbfModel.jl:588 compares with the tolerance of the LESS precise argument type (the default config's
    # delta_max is a Float32 literal, AbstractConfigInterface.jl:31)
    rtol = Float64(Base.rtoldefault(typeof(get_delta(iter_data)), typeof(delta_max(algo_config)), 0))
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
        _drop_kept!(meta)                  # rounds 1-3 came from another group: from-scratch build (mrbf_build) for this one
        empty!(meta.round4_indices)
        if rcfg.optimized_sampling
            Δ_2 = rcfg.θ_enlarge_2 * Δ_max
            lb_2 = Vector{Float64}(max.(lb, x .- Δ_2)); ub_2 = Vector{Float64}(min.(ub, x .+ Δ_2))
            found = Int32.(_collect_indices(meta)); n_found = Int32[length(found)]
            with_ctx(cfg.device) do ctx
                GC.@preserve sites found lb_2 ub_2 r4 begin
                    _check(ctx, ccall((:mrbf_round4, LIBMRBF), Cint,
                        (Ptr{Cvoid}, Ref{MrbfCfg}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                         Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                        ctx, ccfg, 1, n, n_db[1], sites, n_db, lb_2, ub_2, length(found), found, n_found, 0, C_NULL, C_NULL,
                        max_points, r4, n_r4, status))
                end
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
    lbv = Vector{Float64}(lb); ubv = Vector{Float64}(ub)
    old = _kept(meta)
    prepared = Ref{Ptr{Cvoid}}(isnothing(old) ? C_NULL : old.prepared)   # an earlier handle is recycled by the library
    isnothing(old) || (old.prepared = C_NULL)                             # ownership moves through the call
    with_ctx(cfg.device) do ctx
        _check(ctx, ccall((:mrbf_set_isapprox_rtol, LIBMRBF), Cint, (Ptr{Cvoid}, Float64), ctx, rtol))
        GC.@preserve sites x lbv ubv r1 r2 r3 r4 dirs begin
            _check(ctx, ccall((:mrbf_select_points_keep, LIBMRBF), Cint,
                (Ptr{Cvoid}, Ref{MrbfCfg}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
                 Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                ctx, ccfg, 1, n, n_db[1], sites, n_db, xi, x, Δv, Δ_max, lbv, ubv, flags_in, max_new,
                r1, n_r1, r2, n_r2, r3, n_r3, max_points, r4, n_r4, dirs, n_dirs, flags_out, status, prepared))
        end
    end
    _keep!(meta, KeptSelection(prepared[], Int(n_db[1]), Int32(x_index), r1, n_r1, r2, n_r2, r3, n_r3))
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

# prepare_improve_model is n-sized scalar logic on ids and one wall step: the reference method is reused unchanged.
# It may append a site to round1_indices, after which the kept factorisation no longer describes the training set.
function prepare_improve_model(mod::Union{Nothing,GpuRbfModel}, meta::RbfMeta, cfg::GpuRbfConfig, args...; kwargs...)
    n1 = length(meta.round1_indices)
    meta = prepare_improve_model(nothing, meta, cfg.rbf, args...; kwargs...)      # RbfModel.jl:699-732
    length(meta.round1_indices) == n1 || _drop_kept!(meta)
    return meta
end

# ---------------------------------------------------------------------------------------------------------------
# init_model / update_model / improve_model   (RbfModel.jl:738-776)  -> mrbf_build_prepared, else mrbf_build
# ---------------------------------------------------------------------------------------------------------------
init_model(meta::RbfMeta, cfg::GpuRbfConfig, args...; kwargs...) = update_model(nothing, meta, cfg, args...; kwargs...)
improve_model(mod, meta::RbfMeta, cfg::GpuRbfConfig, args...; kwargs...) = update_model(mod, meta, cfg, args...; kwargs...)

# the previous model's device buffers are recycled by the library when the shapes match (the container swaps the new
# model in anyway, SurrogateContainer.jl:376-382): take the handle out of the old object so that its finalizer is a no-op
_take_handle!(mod::GpuRbfModel) = (h = mod.handle; mod.handle = C_NULL; h)
_take_handle!(::Nothing) = C_NULL

function update_model(mod::Union{Nothing,GpuRbfModel}, meta::RbfMeta, cfg::GpuRbfConfig, func_indices, mop, scal, iter_data, sdb, ac; kwargs...)
    db = get_sub_db(sdb, func_indices)
    Δ = Float64(get_delta(iter_data))
    ccfg = Ref(_c_cfg(cfg.rbf, Δ))
    ids = _collect_indices(meta)                                        # centre, r1, r2, r3, r4 -- RbfModel.jl:178-186
    n = length(get_site(db, ids[1])); k = length(get_value(db, ids[1]))
    handle = Ref{Ptr{Cvoid}}(_take_handle!(mod)); status = Int32[0]
    kept = _kept(meta)
    rc = with_ctx(cfg.device) do ctx
        if !isnothing(kept) && kept.prepared != C_NULL && !(cfg.rbf.shape_parameter isa String)
            # round 4 kept its factorisation for exactly this training set: two triangular solves finish the model
            N0 = kept.n_db
            sites = reduce(hcat, (Vector{Float64}(get_site(db, i)) for i in 1:N0))                 # n x N0, the arrays the selection saw
            values = reduce(hcat, (let v = get_value(db, i); isempty(v) ? fill(NaN, k) : Vector{Float64}(v) end for i in 1:N0))   # k x N0
            r3v = zeros(Float64, k, n)                                                           # values of the new round-3 sites
            for (j, rid) in enumerate(meta.round3_indices)
                r3v[:, j] .= get_value(db, rid)
            end
            xi = Int32[kept.x_index]
            GC.@preserve sites values r3v kept begin
                ccall((:mrbf_build_prepared, LIBMRBF), Cint,
                    (Ptr{Cvoid}, Ref{MrbfCfg}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32},
                     Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}, Ptr{Int32}),
                    ctx, ccfg, kept.prepared, k, sites, values, kept.r3, r3v, xi, kept.r1, kept.n_r1, kept.r2, kept.n_r2, kept.n_r3,
                    handle, status)
            end
        else
            sites = reduce(hcat, (Vector{Float64}(get_site(db, i)) for i in ids))      # n x N   (column-major == N x n row-major AoS)
            values = reduce(hcat, (Vector{Float64}(get_value(db, i)) for i in ids))    # k x N
            N = size(sites, 2); Nv = Int32[N]
            GC.@preserve sites values begin
                ccall((:mrbf_build, LIBMRBF), Cint,
                    (Ptr{Cvoid}, Ref{MrbfCfg}, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}, Ptr{Int32}),
                    ctx, ccfg, 1, n, k, N, Nv, sites, values, C_NULL, handle, status)
            end
        end
    end
    # MRBF_ENUMERIC: the handle is valid, status[1] > 0 says the reduced kernel matrix was not positive definite (the reference's
    # `\` would throw SingularException or return garbage on the same duplicated sites); every other error released the handle
    if rc != MRBF_OK && rc != MRBF_ENUMERIC
        with_ctx(ctx -> _check(ctx, rc), cfg.device)
    end
    rc == MRBF_ENUMERIC && @warn "GpuRbfModel: interpolation system not positive definite (duplicated sites?)" status[1]
    return GpuRbfModel(handle[], cfg.device, n, k, meta.fully_linear, status[1]), meta
end

# ---------------------------------------------------------------------------------------------------------------
# eval_models / get_gradient / get_jacobian   (RbfModel.jl:783-800)  -> mrbf_eval
# ---------------------------------------------------------------------------------------------------------------
function _eval(mod::GpuRbfModel, x̂::Vec; values::Bool, jac::Bool)
    x = Vector{Float64}(x̂)
    Y = values ? zeros(Float64, mod.k) : Float64[]
    J = jac ? zeros(Float64, mod.n, mod.k) : Float64[]          # C layout k x n row-major == Julia n x k column-major
    with_ctx(mod.device) do ctx
        GC.@preserve x Y J mod begin
            _check(ctx, ccall((:mrbf_eval, LIBMRBF), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx, mod.handle, 1, x, values ? pointer(Y) : C_NULL, jac ? pointer(J) : C_NULL))
        end
    end
    return Y, J
end
eval_models(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec) = _eval(mod, x̂; values = true, jac = false)[1]
eval_models(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec, ℓ) = eval_models(mod, scal, x̂)[ℓ]       # ℓ::Int or Vector{Int}
get_jacobian(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec, rows = nothing) =
    (J = permutedims(_eval(mod, x̂; values = false, jac = true)[2]); isnothing(rows) ? J : J[rows, :])
get_gradient(mod::GpuRbfModel, scal::AbstractVarScaler, x̂::Vec, ℓ) = vec(get_jacobian(mod, scal, x̂, [ℓ]))

# ---------------------------------------------------------------------------------------------------------------
# Descent hooks.  `compute_descent_step` calls `_backtrack(x_n, d, σ, ω, sc, desc_cfg, scal)` (descent.jl:313) and
# `get_criticality` calls `_steepest_descent_direction(x_n, ∇m, lb, ub, A_eq, b_eq, A_ineq, b_ineq, normalize)` (:239).
# The methods below are more specific than the reference's (typed `sc` / typed arrays), so Julia dispatches to them; they take
# the device path when it applies and `invoke` the reference method otherwise -- no behaviour is lost.
# ---------------------------------------------------------------------------------------------------------------

# all objective outputs of the container come from ONE GpuRbfModel, in model-output order?
function _gpu_objective_model(sc::SurrogateContainer)
    model = nothing; cols = Int[]
    for ind in get_objective_indices(sc)
        s = get_surrogates(sc, ind)
        s isa RefSurrogate || return nothing
        m = s.model_ref[]
        (m isa GpuRbfModel && (isnothing(model) || m === model)) || return nothing
        model = m; append!(cols, s.output_indices)
    end
    (isnothing(model) || cols != collect(1:model.k)) && return nothing
    return model
end

# Armijo backtracking over the surrogate (descent.jl:150-185): one launch for all step sizes instead of <= 118 sequential
# model evaluations; returns the step the sequential loop would have stopped at.
function _backtrack(x::AbstractVector{F}, dir, step_size, ω, sc::SurrogateContainer, cfg::SteepestDescentConfig, scal::AbstractVarScaler) where F<:AbstractFloat
    mod = _gpu_objective_model(sc)
    if isnothing(mod) || F != Float64
        return invoke(_backtrack, Tuple{AbstractVector{F},Any,Any,Any,Any,Any,Any}, x, dir, step_size, ω, sc, cfg, scal)
    end
    n = length(x); k = mod.k
    xv = Vector{Float64}(x); dv = Vector{Float64}(dir)
    idx = Int32[0]; σ = Float64[0]; x₊ = zeros(n); mx = zeros(k); mx₊ = zeros(k)
    s0 = Float64[step_size]; om = Float64[ω]
    min_step = cfg.min_stepsize >= 0 ? cfg.min_stepsize : eps(Float64)
    with_ctx(mod.device) do ctx
        GC.@preserve xv dv s0 om idx σ x₊ mx mx₊ mod begin
            _check(ctx, ccall((:mrbf_backtrack, LIBMRBF), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Int32, Int32,
                 Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx, mod.handle, xv, dv, s0, om, cfg.armijo_const_rhs, cfg.armijo_const_shrink,
                min_step, cfg.max_loops, cfg.strict_backtracking, idx, σ, x₊, mx, mx₊))
        end
    end
    return x₊, mx₊, σ[1] .* dv
end

# Constrained steepest-descent direction (descent.jl:91-135): the LP the reference hands to JuMP + OSQP (eps_rel = 1e-5), solved
# exactly on the device when the MOP has no linear / linearised constraints and at most 8 objectives.
const GPU_DESCENT_LP = Ref(true)        # set to false to keep OSQP in charge
const GPU_DESCENT_DEVICE = Ref(0)
function _steepest_descent_direction(x::Vector{Float64}, ∇F::Matrix{Float64}, lb::Vec, ub::Vec,
        A_eq = [], b_eq = [], A_ineq = [], b_ineq = [], normalize = true)
    k, n = size(∇F)
    if !GPU_DESCENT_LP[] || !isempty(A_eq) || !isempty(A_ineq) || k > 8
        return invoke(_steepest_descent_direction, Tuple{AbstractVector{Float64},Mat,Vec,Vec,Any,Any,Any,Any,Any},
                      x, ∇F, lb, ub, A_eq, b_eq, A_ineq, b_ineq, normalize)
    end
    jac = permutedims(∇F)                       # k x n row-major == n x k column-major
    lbv = Vector{Float64}(lb); ubv = Vector{Float64}(ub)
    d = zeros(n); ω = Float64[0]; iters = Int32[0]; status = Int32[0]
    with_ctx(GPU_DESCENT_DEVICE[]) do ctx
        GC.@preserve jac x lbv ubv d ω iters status begin
            _check(ctx, ccall((:mrbf_descent_direction, LIBMRBF), Cint,
                (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                ctx, 1, n, k, jac, x, lbv, ubv, normalize, d, ω, iters, status))
        end
    end
    status[1] == 0 || return zeros(n), -Inf     # descent.jl:129-133
    return d, ω[1]
end

# ---------------------------------------------------------------------------------------------------------------
# Pascoletti-Serafini inner solves (descent.jl:369-387, 404-412, 478-500).  The reference builds NLopt handles that call the
# surrogates one point at a time (AbstractSurrogateInterface.jl:98-106); with one GPU RBF group holding the objectives (and, behind
# them, the nonlinear inequality constraints of the MOP) and `:GN_ISRES` -- the default -- both NLopt runs become one library call each.
# Hooked in by overloading the two workers `get_criticality(::PascolettiSerafiniConfig, ...)` calls (descent.jl:537, 552); every
# other case (several surrogate groups, linear constraints of the MOP, a polish algorithm, another NLopt algorithm) goes back to NLopt.
# ---------------------------------------------------------------------------------------------------------------
const GPU_PS = Ref(true)
const GPU_PS_SEED = Ref(Int64(0))

function _ps_solve_gpu(mod::GpuRbfModel, x::Vector{Float64}, lb::Vector{Float64}, ub::Vector{Float64}, mx, r, n_obj::Int,
        objective::Int, max_evals::Int)
    n = length(x); k = mod.k
    fmin = Float64[0]; xmin = zeros(n); ymin = zeros(k); found = Int32[0]; used = Int32[0]
    mxv = isnothing(mx) ? Float64[] : Vector{Float64}(mx)
    rv = isnothing(r) ? Float64[] : vcat(Vector{Float64}(r), ones(k - length(r)))
    with_ctx(mod.device) do ctx
        GC.@preserve x lb ub mxv rv fmin xmin ymin found used mod begin
            _check(ctx, ccall((:mrbf_ps_solve, LIBMRBF), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Int32,
                 Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                ctx, mod.handle, x, lb, ub, isnothing(mx) ? C_NULL : pointer(mxv), isnothing(r) ? C_NULL : pointer(rv),
                n_obj, objective, -1, max_evals, GPU_PS_SEED[], fmin, xmin, ymin, found, used))
        end
    end
    GPU_PS_SEED[] += 1
    return fmin[1], xmin, ymin, found[1] == 1
end

# compute_local_ideal_point (descent.jl:404-412): one `_min_component` per objective
function compute_local_ideal_point_gpu(mod::GpuRbfModel, x_n, lb_eff, ub_eff, n_obj::Int, MAX_EVALS::Int)
    x = Vector{Float64}(x_n); lb = Vector{Float64}(lb_eff); ub = Vector{Float64}(ub_eff)
    return [ begin
                 f, _, _, ok = _ps_solve_gpu(mod, x, lb, ub, nothing, nothing, n_obj, l - 1, MAX_EVALS)
                 ok ? f : Inf
             end for l = 1:n_obj ]
end

# _ps_optimization (descent.jl:478-500): returns (tau, x_min, ret) like the reference
function _ps_optimization_gpu(mod::GpuRbfModel, x_n, lb_eff, ub_eff, mx, r, n_obj::Int, MAX_EVALS::Int)
    τ, x_min, _, ok = _ps_solve_gpu(mod, Vector{Float64}(x_n), Vector{Float64}(lb_eff), Vector{Float64}(ub_eff), mx, r, n_obj, -1, MAX_EVALS)
    return τ, x_min, ok ? :MAXEVAL_REACHED : :FAILURE
end

# ---------------------------------------------------------------------------------------------------------------
# Multi-GPU: the final gather of per-instance results (SURVEY §8(e)).  One Julia process / thread per GPU runs its shard of the
# multistart instances with no communication; at the end `gather_results` exchanges the result rows with one ncclAllGather.
# The 128-byte NCCL id travels over whatever channel the host already has (Distributed.jl `remotecall_fetch`, MPI.bcast, a file).
# ---------------------------------------------------------------------------------------------------------------
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:mrbf_comm_unique_id, LIBMRBF), Cint, (Ptr{UInt8},), id)
    rc == MRBF_OK || error("mrbf_comm_unique_id failed ($rc): is libnccl.so.2 loadable (MRBF_NCCL_LIB)?")
    return id
end

mutable struct MrbfComm
    handle::Ptr{Cvoid}; rank::Int; world::Int
end
function MrbfComm(device::Int, id::Vector{UInt8}, rank::Int, world::Int)       # rank is 0-based like NCCL's
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mrbf_comm_init, LIBMRBF), Cint, (Cint, Ptr{UInt8}, Int32, Int32, Ref{Ptr{Cvoid}}), device, id, rank, world, ref)
    rc == MRBF_OK || error("mrbf_comm_init failed ($rc)")
    c = MrbfComm(ref[], rank, world)
    finalizer(cc -> (cc.handle == C_NULL || ccall((:mrbf_comm_destroy, LIBMRBF), Cvoid, (Ptr{Cvoid},), cc.handle); cc.handle = C_NULL), c)
    return c
end

"`rows`: width x count matrix of this rank (one column per instance) -> vector of per-rank matrices."
function gather_results(comm::MrbfComm, rows::Matrix{Float64}, max_count::Int)
    width, count = size(rows)
    all_rows = zeros(Float64, width, max_count, comm.world); counts = zeros(Int32, comm.world)
    rc = GC.@preserve rows all_rows counts ccall((:mrbf_gather, LIBMRBF), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int32}),
        comm.handle, rows, count, width, max_count, all_rows, counts)
    rc == MRBF_OK || error("mrbf_gather: " * unsafe_string(ccall((:mrbf_comm_last_error, LIBMRBF), Cstring, (Ptr{Cvoid},), comm.handle)))
    return [all_rows[:, 1:counts[r], r] for r in 1:comm.world]
end

# user-facing: add_objective!(mop, f; model_cfg = GpuRbfConfig(rbf = RbfConfig(kernel = :multiquadric)), n_out = 2)
