# SSEB200.jl — the reference-side binding of libsse_b200.so.
#
# Adds a third `AbstractParallelism` subtype (next to `Serial` / `Threaded`,
# src/Solvers/Solvers.jl:72,79-80) whose `semi_discrete_residual!` method is one `ccall`.
# Everything else of StableSpectralElements.jl (ConservationLaws, SpatialDiscretization,
# semidiscretize, ODEProblem, Analysis, File) is used unchanged.
#
# NOTE: Julia is not installed in the build image, so this file cannot be executed there.  What IS checked mechanically
# (tests/test_julia_binding.py, CPU suite): every `ccall` below names a symbol include/sse_b200.h declares, with the same
# number of arguments and C-compatible argument / return types, and the two struct mirrors have the header's field order
# and types.  The same entry points are exercised at run time through the Python ctypes mirror
# (cloud.jl_b200/sse_b200/_lib.py), which binds the identical C ABI.
module SSEB200

using StableSpectralElements
using StableSpectralElements.Solvers: AbstractParallelism, Solver, FluxDifferencingOperators,
    ReferenceOperators, PhysicalOperators, StandardForm, FluxDifferencingForm,
    WeightAdjustedSolver, DiagonalSolver, CholeskySolver
using StableSpectralElements.ConservationLaws
using StableSpectralElements.MatrixFreeOperators: WarpedTensorProductMap2D, WarpedTensorProductMap3D
using LinearMaps: UniformScalingMap

const libsse = get(ENV, "SSE_B200_LIB", "libsse_b200.so")

# --- mirror of include/sse_b200.h -----------------------------------------------------------
struct SSEConfig
    abi_version::Int32
    d::Int32
    N_c::Int32; N_p::Int32; N_q::Int32; N_f::Int32; N_fac::Int32
    p::Int32
    N_e::Int64
    N_ghost::Int64
    pde::Int32; form::Int32; inviscid_flux::Int32; viscous_flux::Int32
    two_point_flux::Int32; mass_solver::Int32; v_kind::Int32
    M1d::NTuple{3, Int32}
    half_lambda::Float64
    a::NTuple{3, Float64}
    b::Float64
    gamma::Float64
end

struct SSEArrays
    V::Ptr{Float64}; A::Ptr{Float64}; B::Ptr{Float64}; C::Ptr{Float64}
    sigma_i::Ptr{Int64}; sigma_o::Ptr{Int64}
    R::Ptr{Float64}; W::Ptr{Float64}; Bf::Ptr{Float64}
    D::NTuple{3, Ptr{Float64}}
    S::NTuple{3, Ptr{Float64}}
    Cfd::Ptr{Float64}
    J_q::Ptr{Float64}; Lambda_q::Ptr{Float64}; J_f::Ptr{Float64}; nJf::Ptr{Float64}
    nJq::Ptr{Float64}; nref::Ptr{Float64}
    VOL::Ptr{Float64}; FAC::Ptr{Float64}
    mapP::Ptr{Int64}
end

check(rc::Int32) = rc == 0 ? nothing :
    error("libsse_b200 status $rc: ", unsafe_string(ccall((:sse_last_error_string, libsse), Cstring, ())))

"""
    CUDAB200(device = 0)

Parallelism tag selecting the B200 residual.  Pass it as `parallelism = CUDAB200()` to
`semidiscretize` (Solvers.jl:429-452).
"""
mutable struct CUDAB200 <: AbstractParallelism
    device::Int32
    handle::Ptr{Cvoid}
    CUDAB200(device = 0) = new(Int32(device), C_NULL)
end

"""
Device-resident state: an `AbstractArray{Float64,3}` of size (N_p, N_c, N_e) whose memory lives on the GPU of its handle.
It implements the part of the array interface a time integrator touches between residual calls -- `similar`, `zero`,
`copyto!`, `fill!`, and broadcasts that are linear combinations of states (`@. tmp = A*tmp + dt*k`, `@. u = u + B*tmp`, the
forms OrdinaryDiffEq's low-storage Runge-Kutta methods emit, test/test_driver.jl:77-83) -- on top of `sse_axpby`, so the
state never returns to the host.  Scalar indexing is refused (it would be one PCIe round trip per entry); `Array(x)`
downloads.  Memory is released by a finalizer (`sse_state_free`).
"""
mutable struct DeviceState <: AbstractArray{Float64, 3}
    ptr::Ptr{Float64}
    dims::NTuple{3, Int}
    par::CUDAB200
    function DeviceState(ptr, dims, par)
        x = new(ptr, dims, par)
        finalizer(x) do y
            y.ptr == C_NULL || ccall((:sse_state_free, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}), y.par.handle, y.ptr)
            y.ptr = C_NULL
        end
        return x
    end
end
Base.size(x::DeviceState) = x.dims
Base.getindex(::DeviceState, I...) = error("DeviceState lives on the GPU: use Array(x)")
Base.setindex!(::DeviceState, v, I...) = error("DeviceState lives on the GPU: use copyto!(x, host_array)")
Base.show(io::IO, x::DeviceState) = print(io, "DeviceState", x.dims, " on GPU ", x.par.device)
Base.show(io::IO, ::MIME"text/plain", x::DeviceState) = show(io, x)

function alloc_state(par::CUDAB200, dims)
    p = Ref{Ptr{Float64}}()
    check(ccall((:sse_state_alloc, libsse), Int32, (Ptr{Cvoid}, Ptr{Ptr{Float64}}), par.handle, p))
    return DeviceState(p[], Tuple(dims), par)
end
Base.similar(x::DeviceState) = alloc_state(x.par, x.dims)
Base.similar(x::DeviceState, ::Type{Float64}) = alloc_state(x.par, x.dims)
function Base.similar(x::DeviceState, ::Type{Float64}, dims::Dims{3})
    dims == x.dims || error("a DeviceState has the size of its Solver: ", x.dims)
    return alloc_state(x.par, dims)
end
Base.fill!(x::DeviceState, v::Real) =
    (check(ccall((:sse_state_fill, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64), x.par.handle, x.ptr, Float64(v))); x)
Base.zero(x::DeviceState) = fill!(similar(x), 0.0)
upload!(x::DeviceState, h::Array{Float64, 3}) =
    (check(ccall((:sse_state_upload, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), x.par.handle, x.ptr, h)); x)
function Base.Array(x::DeviceState)
    h = Array{Float64}(undef, x.dims)
    check(ccall((:sse_state_download, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), x.par.handle, h, x.ptr))
    return h
end
# y = a x + b y on the device
axpby!(a, x::DeviceState, b, y::DeviceState) =
    (check(ccall((:sse_axpby, libsse), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Float64, Ptr{Float64}),
        x.par.handle, Float64(a), x.ptr, Float64(b), y.ptr)); y)
Base.copyto!(dst::DeviceState, src::DeviceState) = dst.ptr == src.ptr ? dst : axpby!(1.0, src, 0.0, dst)
Base.copyto!(dst::DeviceState, src::Array{Float64, 3}) = upload!(dst, src)
Base.copyto!(dst::Array{Float64, 3}, src::DeviceState) = copyto!(dst, Array(src))
Base.copy(x::DeviceState) = copyto!(similar(x), x)
device_synchronize(x::DeviceState) = check(ccall((:sse_synchronize, libsse), Int32, (Ptr{Cvoid},), x.par.handle))

# --- broadcasting: linear combinations of DeviceStates with scalar coefficients are lowered to sse_axpby ------------------
struct DeviceStyle <: Broadcast.AbstractArrayStyle{3} end
DeviceStyle(::Val{3}) = DeviceStyle()
DeviceStyle(::Val{N}) where {N} = Broadcast.DefaultArrayStyle{N}()
Base.BroadcastStyle(::Type{DeviceState}) = DeviceStyle()
Base.similar(bc::Broadcast.Broadcasted{DeviceStyle}, ::Type{Float64}) = similar(first_state(bc))
first_state(x::DeviceState) = x
first_state(x) = nothing
function first_state(bc::Broadcast.Broadcasted)
    for a in bc.args
        s = first_state(a)
        s === nothing || return s
    end
    return nothing
end
# a broadcast expression as a list of (coefficient, state) terms; anything that is not linear in the states is refused
const Term = Tuple{Float64, DeviceState}
lin(x::DeviceState) = Term[(1.0, x)]
lin(x::Ref) = lin(x[])
lin(x) = error("unsupported operand in a DeviceState broadcast: ", typeof(x), " (only linear combinations of states)")
scalar(x::Number) = Float64(x)
scalar(x::Ref{<:Number}) = Float64(x[])
scalar(x) = nothing
function lin(bc::Broadcast.Broadcasted)
    f, args = bc.f, bc.args
    if f === (+)
        return reduce(vcat, map(lin, args))
    elseif f === (-) && length(args) == 1
        return Term[(-c, x) for (c, x) in lin(args[1])]
    elseif f === (-) && length(args) == 2
        return vcat(lin(args[1]), Term[(-c, x) for (c, x) in lin(args[2])])
    elseif f === (*)
        k, rest = 1.0, Any[]
        for a in args
            s = scalar(a)
            s === nothing ? push!(rest, a) : (k *= s)
        end
        length(rest) == 1 || error("a DeviceState broadcast may multiply a state by scalars only")
        return Term[(k * c, x) for (c, x) in lin(rest[1])]
    elseif f === (/) && length(args) == 2 && scalar(args[2]) !== nothing
        return Term[(c / scalar(args[2]), x) for (c, x) in lin(args[1])]
    elseif f === muladd && length(args) == 3          # muladd(a, x, y) = a x + y   (what @muladd / @.. emit)
        s1, s2 = scalar(args[1]), scalar(args[2])
        prod = s1 !== nothing ? Term[(s1 * c, x) for (c, x) in lin(args[2])] :
               s2 !== nothing ? Term[(s2 * c, x) for (c, x) in lin(args[1])] :
               error("muladd of two states is not linear")
        return vcat(prod, lin(args[3]))
    elseif f === identity && length(args) == 1
        return lin(args[1])
    end
    error("unsupported function in a DeviceState broadcast: ", f)
end
function Base.copyto!(dest::DeviceState, bc::Broadcast.Broadcasted{DeviceStyle})
    terms = lin(bc)                                              # nested Broadcasted objects are walked by `lin` itself
    # merge the coefficients of repeated states; the destination's own term becomes the `b` of the first axpby
    coef = Dict{Ptr{Float64}, Term}()
    for (c, x) in terms
        coef[x.ptr] = haskey(coef, x.ptr) ? (coef[x.ptr][1] + c, x) : (c, x)
    end
    b = haskey(coef, dest.ptr) ? coef[dest.ptr][1] : 0.0
    delete!(coef, dest.ptr)
    if isempty(coef)
        return axpby!(b, dest, 0.0, dest)                        # dest .= b .* dest
    end
    for (c, x) in values(coef)
        axpby!(c, x, b, dest)                                   # dest = c x + b dest, then accumulate with b = 1
        b = 1.0
    end
    return dest
end
Base.copy(bc::Broadcast.Broadcasted{DeviceStyle}) = copyto!(similar(bc, Float64), bc)

pde_id(::LinearAdvectionEquation) = Int32(0)
pde_id(::LinearAdvectionDiffusionEquation) = Int32(1)
pde_id(::EulerEquations) = Int32(2)
pde_id(::InviscidBurgersEquation) = Int32(3)
pde_id(::ViscousBurgersEquation) = Int32(4)
flux_id(::LaxFriedrichsNumericalFlux) = Int32(0)
flux_id(::CentralNumericalFlux) = Int32(1)
flux_id(::EntropyConservativeNumericalFlux) = Int32(2)
halfλ(f::LaxFriedrichsNumericalFlux) = f.halfλ
halfλ(_) = 0.0
two_point_id(::ConservativeFlux) = Int32(0)
two_point_id(::EntropyConservativeFlux) = Int32(1)
mass_id(::WeightAdjustedSolver) = Int32(0)
mass_id(::DiagonalSolver) = Int32(1)
mass_id(::CholeskySolver) = Int32(2)          # the library factorises V' WJ_k V itself (mass_matrix.jl:30-39)

"""
    attach!(solver::Solver{<:Any,<:Any,<:Any,<:Any,CUDAB200}, spatial_discretization)

Uploads the Solver's operators and geometric factors (the arrays the reference constructors
at Solvers.jl:287-376 / operators.jl:1-160 hold) and stores the opaque handle in the tag.
"""
function attach!(solver::Solver, sd::SpatialDiscretization{d}; n_ghost::Integer = 0) where {d}
    par = solver.parallelism::CUDAB200
    law, ops, form = solver.conservation_law, solver.operators, solver.form
    ra, gf = sd.reference_approximation, sd.geometric_factors
    N_p, N_c, N_e = size(solver)
    keep = Any[]                                    # GC roots for the duration of sse_create
    ptr(x::Array) = (push!(keep, x); pointer(x))
    dense(L) = ptr(Matrix(L))
    nul = Ptr{Float64}(C_NULL)
    V = ra.V
    v_kind, pV, pA, pB, pC, pσi, pσo, M1d = if V isa UniformScalingMap
        Int32(0), nul, nul, nul, nul, Ptr{Int64}(C_NULL), Ptr{Int64}(C_NULL), (Int32(0), Int32(0), Int32(0))
    elseif V isa WarpedTensorProductMap3D
        Int32(2), dense(V), ptr(Array(V.A)), ptr(Array(V.B)), ptr(Array(V.C)), ptr(Array{Int64}(V.σᵢ)),
        ptr(Array{Int64}(V.σₒ)), Int32.((size(V.A, 1), size(V.B, 1), size(V.C, 1)))
    elseif V isa WarpedTensorProductMap2D
        Int32(2), dense(V), ptr(Array(V.A)), ptr(Array(V.B)), nul, ptr(Array{Int64}(V.σᵢ)),
        ptr(Array{Int64}(V.σₒ)), Int32.((size(V.A, 1), size(V.B, 1), 0))
    else
        Int32(1), dense(V), nul, nul, nul, Ptr{Int64}(C_NULL), Ptr{Int64}(C_NULL), (Int32(0), Int32(0), Int32(0))
    end
    form_id, pD, pS, pC_fd, pΛ, pVOL, pFAC = if ops isa FluxDifferencingOperators
        Int32(2), (nul, nul, nul), ntuple(m -> m <= d ? dense(ops.S[m]) : nul, 3),
        isnothing(ops.C) ? nul : dense(ops.C), ptr(ops.Λ_q), nul, nul
    elseif ops isa ReferenceOperators
        Λη = apply_reference_mapping(gf, ra.reference_mapping).Λ_q
        Int32(0), ntuple(m -> m <= d ? dense(ops.D[m]) : nul, 3), (nul, nul, nul), nul, ptr(Λη), nul, nul
    else  # PhysicalOperators: VOL (N_p,N_q,d,N_e), FAC (N_p,N_f,N_e)
        VOL = Array{Float64}(undef, N_p, ra.N_q, d, N_e)
        FAC = Array{Float64}(undef, N_p, ra.N_f, N_e)
        for k in 1:N_e
            for m in 1:d
                VOL[:, :, m, k] .= Matrix(ops.VOL[k][m])
            end
            FAC[:, :, k] .= Matrix(ops.FAC[k])
        end
        Int32(1), (nul, nul, nul), (nul, nul, nul), nul, ptr(gf.Λ_q), ptr(VOL), ptr(FAC)
    end
    nfac = length(ra.reference_element.fv)
    npf = ra.N_f ÷ nfac
    nref = [ra.reference_element.nrstJ[m][npf * (f - 1) + 1] for m in 1:d, f in 1:nfac]
    tp = form isa FluxDifferencingForm ? two_point_id(form.two_point_flux) : Int32(0)
    a = ntuple(m -> (hasproperty(law, :a) && m <= d) ? Float64(law.a[m]) : 0.0, 3)
    cfg = SSEConfig(1, d, N_c, N_p, ra.N_q, ra.N_f, nfac, ra.approx_type.p, N_e, Int64(n_ghost),
        pde_id(law), form_id, flux_id(form.inviscid_numerical_flux),
        (law isa LinearAdvectionDiffusionEquation || law isa ViscousBurgersEquation) ? Int32(1) : Int32(0), tp, mass_id(solver.mass_solver), v_kind, M1d,
        halfλ(form.inviscid_numerical_flux), a, hasproperty(law, :b) ? law.b : 0.0,
        hasproperty(law, :γ) ? law.γ : 1.4)
    arr = SSEArrays(pV, pA, pB, pC, pσi, pσo, dense(ra.R), ptr(Vector(ra.W.diag)), ptr(Vector(ra.B.diag)),
        pD, pS, pC_fd, ptr(gf.J_q), pΛ, ptr(gf.J_f), ptr(gf.nJf), ptr(gf.nJq), ptr(nref), pVOL, pFAC,
        ptr(Array{Int64}(solver.connectivity)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:sse_create, libsse), Int32,
        (Ref{SSEConfig}, Ref{SSEArrays}, Int32, Ptr{Ptr{Cvoid}}), cfg, arr, par.device, h))
    par.handle = h[]
    finalizer(p -> ccall((:sse_destroy, libsse), Int32, (Ptr{Cvoid},), p.handle), par)
    return solver
end

# The drop-in: same signature as Solvers.jl:474-483, dispatched on the new parallelism tag.
function StableSpectralElements.Solvers.semi_discrete_residual!(dudt::DeviceState, u::DeviceState,
        solver::Solver{<:Any, <:Any, <:Any, <:Any, CUDAB200}, t::Float64 = 0.0)
    check(ccall((:sse_rhs, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64),
        solver.parallelism.handle, u.ptr, dudt.ptr, t))
    return dudt
end

# Host arrays (OrdinaryDiffEq with CPU state, the save callback's `similar(integrator.u)`, File/save.jl:58-61): one
# ccall of sse_rhs_host, which pipelines upload, both passes and download over element ranges inside the library.
# Page-lock long-lived arrays once with `pin!` (cudaHostRegister) so the copies overlap in both PCIe directions.
pin!(x::Array{Float64}) = (check(ccall((:sse_host_pin, libsse), Int32, (Ptr{Cvoid}, Int64), x, sizeof(x))); x)
unpin!(x::Array{Float64}) = (check(ccall((:sse_host_unpin, libsse), Int32, (Ptr{Cvoid},), x)); x)

function StableSpectralElements.Solvers.semi_discrete_residual!(dudt::Array{Float64, 3}, u::Array{Float64, 3},
        solver::Solver{<:Any, <:Any, <:Any, <:Any, CUDAB200}, t::Float64 = 0.0)
    GC.@preserve dudt u check(ccall((:sse_rhs_host, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Int32),
        solver.parallelism.handle, u, dudt, t, 0))
    return dudt
end

# --- device-resident time integration without the broadcast machinery ----------------------------------------------------
"""
    solve_ck54!(u, solver, tspan, dt; callback = nothing)

`solve(ode, CarpenterKennedy2N54(); adaptive = false, dt)` (test/test_driver.jl:77-83) with the whole step on the device:
one `sse_step_ck54` per step (five fused residual + 2N-storage stages).  `callback(u, t)` runs between steps.
"""
function solve_ck54!(u::DeviceState, solver::Solver{<:Any, <:Any, <:Any, <:Any, CUDAB200}, tspan, dt; callback = nothing)
    tmp, dudt = zero(u), similar(u)
    t, T = Float64(tspan[1]), Float64(tspan[2])
    while t < T - 1e-12 * abs(T)
        h = min(dt, T - t)
        check(ccall((:sse_step_ck54, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64),
            solver.parallelism.handle, u.ptr, tmp.ptr, dudt.ptr, t, h))
        t += h
        callback === nothing || callback(u, t)
    end
    device_synchronize(u)
    return u
end

"conservation / energy / entropy residuals on the device (Analysis/conservation.jl:145-189): N_c + 2 numbers"
function functionals(u::DeviceState, dudt::DeviceState, solver::Solver{<:Any, <:Any, <:Any, <:Any, CUDAB200})
    out = zeros(Float64, size(solver)[2] + 2)
    check(ccall((:sse_functionals, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        solver.parallelism.handle, u.ptr, dudt.ptr, out))
    return out
end

# --- several GPUs driven by this one Julia process (include/sse_b200.h, "multi-GPU") ---------------------------------------
# One Solver per partition (built from the partition's mesh: elements ordered interior first, `mapP` in local + ghost
# numbering, `N_ghost` set in `attach!` through `n_ghost`), one GPU each.
"""
    partition(mapP, N_f, owner, n_parts, rank) -> NamedTuple

Local view of `rank` (0-based) of an element partition, from the mesh's global connectivity `mesh.mapP` (N_f x N_e, 1-based;
Solvers.jl:207) and the owner rank of every element: `elem_gid` (1-based global ids, interior elements first), `mapP` in
local + ghost numbering, `n_interior`, `n_ghost`, and the halo plan (`nbr_rank`, `send_count`, `recv_count`, `send_index`).
Build the partition's `SpatialDiscretization` from the elements `elem_gid`, pass `n_ghost` to `attach!` and the plan to
`halo_plan!`.  Host only; no device is touched.
"""
function partition(mapP::Array{Int64}, N_f::Integer, owner::Vector{Int32}, n_parts::Integer, rank::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sse_partition_create, libsse), Int32, (Ptr{Int64}, Int64, Int32, Ptr{Int32}, Int32, Int32, Ptr{Ptr{Cvoid}}),
        mapP, Int64(length(owner)), Int32(N_f), owner, Int32(n_parts), Int32(rank), h))
    nl, ni, ng, ns, nn = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int32}(0)
    check(ccall((:sse_partition_sizes, libsse), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}),
        h[], nl, ni, ng, nn, ns))
    elem_gid, mapP_local = Vector{Int64}(undef, nl[]), Matrix{Int64}(undef, N_f, nl[])
    nbr_rank, send_count, recv_count = Vector{Int32}(undef, nn[]), Vector{Int64}(undef, nn[]), Vector{Int64}(undef, nn[])
    send_index = Vector{Int64}(undef, ns[])
    check(ccall((:sse_partition_fill, libsse), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        h[], elem_gid, mapP_local, nbr_rank, send_count, recv_count, send_index))
    check(ccall((:sse_partition_destroy, libsse), Int32, (Ptr{Cvoid},), h[]))
    return (; elem_gid, mapP = mapP_local, n_interior = ni[], n_ghost = ng[], nbr_rank, send_count, recv_count, send_index)
end

"join the handles of `pars` into one NCCL communicator (ncclCommInitAll)"
comm_init_all!(pars::Vector{CUDAB200}) =
    check(ccall((:sse_comm_init_all, libsse), Int32, (Ptr{Ptr{Cvoid}}, Int32), [p.handle for p in pars], Int32(length(pars))))
"the same without NCCL: halos by peer-to-peer copies"
comm_init_local!(pars::Vector{CUDAB200}) =
    check(ccall((:sse_comm_init_local, libsse), Int32, (Ptr{Ptr{Cvoid}}, Int32), [p.handle for p in pars], Int32(length(pars))))
"halo plan of one partition: neighbour ranks (0-based), facet nodes sent / received per neighbour, 1-based send list, interior element count"
halo_plan!(par::CUDAB200, nbr_rank::Vector{Int32}, send_count::Vector{Int64}, recv_count::Vector{Int64},
    send_index::Vector{Int64}, n_interior::Integer) =
    check(ccall((:sse_halo_plan, libsse), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64),
        par.handle, Int32(length(nbr_rank)), nbr_rank, send_count, recv_count, send_index, Int64(n_interior)))
"semi_discrete_residual! on all partitions (one NCCL group for the halos of all GPUs)"
function rhs_multi!(dudts::Vector{DeviceState}, us::Vector{DeviceState}, pars::Vector{CUDAB200}, t::Float64 = 0.0)
    check(ccall((:sse_rhs_multi, libsse), Int32, (Ptr{Ptr{Cvoid}}, Int32, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Float64),
        [p.handle for p in pars], Int32(length(pars)), [u.ptr for u in us], [d.ptr for d in dudts], t))
    return dudts
end
"one CarpenterKennedy2N54 step on all partitions"
step_ck54_multi!(us::Vector{DeviceState}, tmps::Vector{DeviceState}, dudts::Vector{DeviceState}, pars::Vector{CUDAB200}, t, dt) =
    check(ccall((:sse_step_ck54_multi, libsse), Int32,
        (Ptr{Ptr{Cvoid}}, Int32, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Float64, Float64),
        [p.handle for p in pars], Int32(length(pars)), [u.ptr for u in us], [x.ptr for x in tmps], [d.ptr for d in dudts],
        Float64(t), Float64(dt)))

end # module
