# IntfAdvB200Ext.jl -- package extension that routes InterfaceAdvection.jl's VOF + CMOM advection path to
# libifadv_b200.so (hand-written sm_100a CUDA, C ABI in include/ifadv.h) for CuArray arguments.
#
# It replaces ext/IntfAdvCUDAExt.jl ON THIS PATH: instead of only allowing scalar indexing (the 8 lines the stock
# extension consists of, ext/IntfAdvCUDAExt.jl:19-23) it overrides the two sweep drivers
#     advectVOF!     (src/advection.jl:34)      and      advectVOFρuu!  (src/flow.jl:165)
# plus the secondary seams u2ρu!/ρu2u! (src/VOFutil.jl:198-211), MPCFL (src/flow.jl:262) and the explicit forcing viscSurfTenρu! /
# updateU! / updateL! (src/flow.jl:113-117,244-259) for CuArray{Float32|Float64}.
# No KernelAbstractions, no multi-backend dispatch, no CPU fallback: unsupported argument combinations raise.
#
# STATUS: UNTESTED.  No Julia toolchain exists in the environment this library was built in, so this file has been written
# against the C header (the ccall signatures are checked against include/ifadv.h by tests/test_abi_and_host.py, which parses
# both) but never executed (INTEGRATION.md).  Install: copy to ext/, add to Project.toml
#     [extensions]  IntfAdvB200Ext = "CUDA"      (instead of IntfAdvCUDAExt)
# and point ENV["IFADV_B200_LIB"] at libifadv_b200.so.
module IntfAdvB200Ext

using CUDA
using InterfaceAdvection
import Random
import InterfaceAdvection: advectVOF!, advectVOFρuu!, u2ρu!, ρu2u!, MPCFL, _scalar_op, cVOF
import InterfaceAdvection: viscSurfTenρu!, updateU!, updateL!
import InterfaceAdvection: psolver!, myproject!
import InterfaceAdvection: LevelSet, redistaning!, computeL!, _redistaningStage!
import InterfaceAdvection: getInterfaceNormal_WH!, getInterfaceNormal_WY!, getInterfaceNormal_Column!, getInterfaceNormal_PCD!,
                           getInterfaceNormal_SLIC!, getInterfaceNormal_MYC!, getInterfaceNormal_Y!, getInterfaceNormal_CD!,
                           getInterfaceNormal_XYLIC!
import InterfaceAdvection: upwind, minmod, Koren, vanAlbada1, Sweby, superbee, TVDcen, TVDdown
import WaterLily
import WaterLily: Flow, quick, vanLeer, cds

const LIB = get(ENV, "IFADV_B200_LIB", "libifadv_b200.so")

__init__() = @assert CUDA.functional()

# the stock extension's only job stays available for the diagnostics that still index scalars
_scalar_op(op::F, ::CUDA.CuArray) where {F<:Function} = CUDA.@allowscalar op()

# ---- function identity -> enum (include/ifadv.h) -------------------------------------------------------------------
const NORMAL_ENUM = IdDict{Any,Cint}(
    getInterfaceNormal_WH! => 0, getInterfaceNormal_WY! => 1, getInterfaceNormal_Column! => 2, getInterfaceNormal_PCD! => 3,
    getInterfaceNormal_SLIC! => 4, getInterfaceNormal_MYC! => 5, getInterfaceNormal_Y! => 6, getInterfaceNormal_CD! => 7,
    getInterfaceNormal_XYLIC! => 8)
const LIMITER_ENUM = IdDict{Any,Cint}(
    upwind => 0, minmod => 1, Koren => 2, vanAlbada1 => 3, Sweby => 4, superbee => 5, TVDcen => 6, TVDdown => 7,
    quick => 8, vanLeer => 9, cds => 10)
normal_enum(f) = get(NORMAL_ENUM, f) do; error("IntfAdvB200Ext: normalScheme $f has no sm_100a kernel (no fallback)"); end
limiter_enum(f) = get(LIMITER_ENUM, f) do; error("IntfAdvB200Ext: limiter $f has no sm_100a kernel (no fallback)"); end
perdir_mask(perdir) = reduce(|, (UInt32(1) << (j - 1) for j in perdir); init=UInt32(0))
dtype_enum(::Type{Float32}) = Cint(0)
dtype_enum(::Type{Float64}) = Cint(1)

struct IfadvReport
    maxf::Cdouble; minf::Cdouble
    argmax::NTuple{3,Int64}; argmin::NTuple{3,Int64}
    dir::Cint; status::Cint
    div_u0::Cdouble; div_u::Cdouble
end
IfadvReport() = IfadvReport(0, 0, (0, 0, 0), (0, 0, 0), 0, 0, 0, 0)

# ---- one context per (device, size, eltype) ----------------------------------------------------------------------------
const CONTEXTS = Dict{Tuple{Int,NTuple{3,Int64},DataType},Ptr{Cvoid}}()
function context(f::CuArray{T,D}) where {T,D}
    Ng = ntuple(i -> i <= D ? Int64(size(f, i)) : Int64(1), 3)
    key = (CUDA.deviceid(CUDA.device()), Ng, T)
    get!(CONTEXTS, key) do
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ifadv_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ref{NTuple{3,Int64}}, Cint, Cint),
                   ctx, D, Ref(Ng), dtype_enum(T), key[1])
        rc == 0 || error("ifadv_create failed ($rc)")
        ctx[]
    end
end
stream_ptr() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
dptr(a::CuArray) = reinterpret(Ptr{Cvoid}, UInt(pointer(a)))
function check(ctx, rc)
    rc == -1 && error("NaN!")                                   # error("NaN!"), src/advection.jl:148
    rc == -5 && error(unsafe_string(ccall((:ifadv_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))  # "divergence, …, is exploding!" :160,180
    rc < 0 && error("ifadv error $rc: " * unsafe_string(ccall((:ifadv_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))
    rc
end
function report(rc, rep::IfadvReport)
    rc > 0 || return
    which, Δ, idx = (rc & 1) != 0 ? ("max", rep.maxf - 1, rep.argmax) : ("min", -rep.minf, rep.argmin)
    println("|∇⋅u⁰| = $(rep.div_u0), |∇⋅u| = $(rep.div_u)")   # reportFillError's diagnostics (advection.jl:151,170)
    Base.printstyled("ERROR: "; color=:red, bold=true)   # printed, not thrown -- like reportFillError (advection.jl:161-166)
    println("$which VOF @ $idx ∉ [0,1] @ direction $(rep.dir), Δf = $Δ")
end
const CHECK_EVERY = Ref(1)   # set to k > 1 to fetch the fill-error report (a stream sync) only every k-th call
const CALLS = Ref(0)

# ---- advectVOF!  (src/advection.jl:34-78) ----------------------------------------------------------------------------------
function advectVOF!(f::CuArray{T,D}, fᶠ, α, n̂, u, u⁰, Δt, c̄, ρuf, λρ, normalScheme; perdir=(), dirO=nothing) where {T<:Union{Float32,Float64},D}
    ctx = context(f)
    dO = isnothing(dirO) ? Random.shuffle(1:D) : dirO
    dirv = Cint[dO...; zeros(Cint, 3 - D)]
    rep = Ref(IfadvReport())
    want = (CALLS[] += 1) % CHECK_EVERY[] == 0
    rc = ccall((:ifadv_advect_vof, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid},
                Cdouble, Cint, Cuint, Ptr{Cint}, Cint, Ptr{IfadvReport}),
               ctx, stream_ptr(), dptr(f), dptr(fᶠ), dptr(α), dptr(n̂), dptr(u), dptr(u⁰), Δt, dptr(c̄), dptr(ρuf),
               λρ, normal_enum(normalScheme), perdir_mask(perdir), dirv, 0, want ? rep : C_NULL)
    report(check(ctx, rc), rep[])
    nothing
end

# ---- advectVOFρuu!  (src/flow.jl:165-210) ------------------------------------------------------------------------------------
function advectVOFρuu!(f::CuArray{T,D}, fᶠ, α, n̂, u, u⁰, Δt, c̄, ρu, r, Φ, ρuf, uStar, uOld, dilaU, dρ, λρ, λ, normalScheme, uBC;
                       perdir=(), exitBC=false, dirO=nothing) where {T<:Union{Float32,Float64},D}
    uBC isa Function && error("IntfAdvB200Ext: function-valued uBC is not supported (no fallback)")
    ctx = context(f)
    dO = isnothing(dirO) ? Random.shuffle(1:D) : dirO
    dirv = Cint[dO...; zeros(Cint, 3 - D)]
    A = Cdouble[uBC...; zeros(3 - D)]
    rep = Ref(IfadvReport())
    want = (CALLS[] += 1) % CHECK_EVERY[] == 0
    rc = ccall((:ifadv_advect_vof_rhouu, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid},
                Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint,
                Ptr{Cdouble}, Cuint, Cint, Ptr{Cint}, Ptr{IfadvReport}),
               ctx, stream_ptr(), dptr(f), dptr(fᶠ), dptr(α), dptr(n̂), dptr(u), dptr(u⁰), Δt, dptr(c̄),
               dptr(ρu), dptr(r), dptr(Φ), dptr(ρuf), dptr(uStar), dptr(uOld), dptr(dilaU), dptr(dρ), λρ, limiter_enum(λ),
               normal_enum(normalScheme), A, perdir_mask(perdir), exitBC ? 1 : 0, dirv, want ? rep : C_NULL)
    report(check(ctx, rc), rep[])
    nothing
end

# ---- secondary seams -------------------------------------------------------------------------------------------------------------
function u2ρu!(ρu::CuArray{T}, u, f::CuArray{T}, λρ) where {T<:Union{Float32,Float64}}   # src/VOFutil.jl:208
    ctx = context(f)
    check(ctx, ccall((:ifadv_u2rhou, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
                     ctx, stream_ptr(), dptr(ρu), dptr(u), dptr(f), λρ)); nothing
end
function ρu2u!(u::CuArray{T}, ρu, f::CuArray{T}, λρ) where {T<:Union{Float32,Float64}}   # src/VOFutil.jl:198
    ctx = context(f)
    check(ctx, ccall((:ifadv_rhou2u, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
                     ctx, stream_ptr(), dptr(u), dptr(ρu), dptr(f), λρ)); nothing
end
function MPCFL(a::Flow{D,T}, c::cVOF; Δt_max=one(T), safetyMargin=T(0.8)) where {D,T<:Union{Float32,Float64}}   # src/flow.jl:262
    a.u isa CuArray || return invoke(MPCFL, Tuple{Flow,cVOF}, a, c; Δt_max, safetyMargin)
    ctx = context(c.f)
    g2 = isnothing(a.g) ? 0.0 : sqrt(sum(abs2, (a.g(i, zeros(T, D), sum(a.Δt)) for i in 1:D)))
    out = Ref{Cdouble}(0)
    check(ctx, ccall((:ifadv_mpcfl, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Ref{Cdouble}),
                     ctx, stream_ptr(), dptr(a.u), a.ν, isnothing(c.μ) ? 0.0 : c.μ, c.λμ, c.λρ, isnothing(c.η) ? 0.0 : c.η, g2,
                     Δt_max, safetyMargin, out))
    T(out[])
end

# ---- explicit forcing between advection and projection (src/flow.jl:113-117, 244-259) --------------------------------------------
function viscSurfTenρu!(r::CuArray{T}, u, Φ, f::CuArray{T}, α, n̂, fbuffer, λμ, μ, λρ, η; perdir=()) where {T<:Union{Float32,Float64}}
    ctx = context(f)
    check(ctx, ccall((:ifadv_visc_surften_rhou, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble,
                      Cdouble, Cdouble, Cuint),
                     ctx, stream_ptr(), dptr(r), dptr(u), dptr(Φ), dptr(f), dptr(α), dptr(n̂), dptr(fbuffer), λμ, isnothing(μ) ? 0.0 : μ, λρ,
                     isnothing(η) ? 0.0 : η, perdir_mask(perdir))); nothing
end
function updateU!(u::CuArray{T}, ρu, ρu⁰, forcing, dt, f::CuArray{T,D}, λρ, tNow, g, uBC, w=one(T)) where {T<:Union{Float32,Float64},D}
    uBC isa Function && error("IntfAdvB200Ext: function-valued uBC is not supported (no fallback)")
    ctx = context(f)
    gv = isnothing(g) ? Ptr{Cdouble}(C_NULL) : Cdouble[(g(i, zeros(T, D), tNow) for i in 1:D)...; zeros(3 - D)]   # constant gravity only
    GC.@preserve gv check(ctx, ccall((:ifadv_update_u, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Cdouble),
                     ctx, stream_ptr(), dptr(u), dptr(ρu), dptr(ρu⁰), dptr(forcing), dt, dptr(f), λρ, gv, w)); nothing
end
function updateL!(μ₀::CuArray{T}, f::CuArray{T}, λρ; perdir=()) where {T<:Union{Float32,Float64}}
    ctx = context(f)
    check(ctx, ccall((:ifadv_update_l, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cuint, Cint),
                     ctx, stream_ptr(), dptr(μ₀), dptr(f), λρ, perdir_mask(perdir), 0)); nothing
end

# ---- pressure projection on WaterLily's Poisson (src/flow.jl:300-347): update!, psolver!, myproject! -------------------------------
# b.L ≡ a.μ₀, b.x ≡ a.p, b.z ≡ a.σ; MultiLevelPoisson: next block.
const CuPoisson{T} = WaterLily.Poisson{T,<:CuArray{T},<:CuArray{T}}
function WaterLily.update!(b::CuPoisson{T}) where {T<:Union{Float32,Float64}}
    ctx = context(b.x)
    check(ctx, ccall((:ifadv_poisson_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                     ctx, stream_ptr(), dptr(b.D), dptr(b.iD), dptr(b.L))); nothing
end
function psolver!(b::CuPoisson{T}; log=false, tol=50eps(T), itmx=6e3) where {T<:Union{Float32,Float64}}
    ctx = context(b.x)
    it = Ref{Cint}(0); r₂ = Ref{Cdouble}(0)
    check(ctx, ccall((:ifadv_psolver, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Cdouble, Cint,
                      Ptr{Cint}, Ptr{Cdouble}),
                     ctx, stream_ptr(), dptr(b.x), dptr(b.ϵ), dptr(b.r), dptr(b.z), dptr(b.L), dptr(b.D), dptr(b.iD), perdir_mask(b.perdir),
                     Cdouble(tol), Cint(itmx), it, r₂))
    push!(b.n, it[]); nothing
end
function myproject!(a::Flow{n,T}, b::CuPoisson{T}, w=1) where {n,T<:Union{Float32,Float64}}
    ctx = context(b.x)
    it = Ref{Cint}(0); r₂ = Ref{Cdouble}(0)
    check(ctx, ccall((:ifadv_myproject, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble,
                      Cuint, Ptr{Cint}, Ptr{Cdouble}),
                     ctx, stream_ptr(), dptr(a.u), dptr(b.x), dptr(b.ϵ), dptr(b.r), dptr(b.z), dptr(b.L), dptr(b.D), dptr(b.iD),
                     Cdouble(T(w) * last(a.Δt)), perdir_mask(b.perdir), it, r₂))
    push!(b.n, it[]); nothing
end

# ---- WaterLily.MultiLevelPoisson (WaterLily's default psolver; inproject!'s second method, src/flow.jl:343-347) -----------------------
# The library keeps the coarse levels (and level 1's D, iD, ϵ, r) in a handle built from ml.x ≡ a.p, ml.L ≡ a.μ₀, ml.z ≡ a.σ; one handle per
# ml object, destroyed with it.  ml.levels[2:end] of the Julia object are not used by the methods below.
const CuMLPoisson{T} = WaterLily.MultiLevelPoisson{T,<:CuArray{T},<:CuArray{T}}
const ML_HANDLES = IdDict{Any,Ptr{Cvoid}}()
function ml_handle(ml::CuMLPoisson{T}) where {T<:Union{Float32,Float64}}
    get!(ML_HANDLES, ml) do
        ctx = context(ml.x)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx, ccall((:ifadv_ml_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Cint),
                         ctx, h, stream_ptr(), dptr(ml.x), dptr(ml.L), dptr(ml.z), perdir_mask(ml.perdir), Cint(10)))
        finalizer(ml) do m
            ctxml = pop!(ML_HANDLES, m, C_NULL)
            ctxml == C_NULL || ccall((:ifadv_ml_destroy, LIB), Cint, (Ptr{Cvoid},), ctxml)
        end
        h[]
    end
end
function WaterLily.update!(ml::CuMLPoisson{T}) where {T<:Union{Float32,Float64}}
    ctxml = ml_handle(ml)
    check(context(ml.x), ccall((:ifadv_ml_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctxml, stream_ptr())); nothing
end
function WaterLily.solver!(ml::CuMLPoisson{T}; log=false, tol=1e-4, itmx=32) where {T<:Union{Float32,Float64}}
    n = Ref{Cint}(0); r₂ = Ref{Cdouble}(0)
    ctxml = ml_handle(ml)
    check(context(ml.x), ccall((:ifadv_ml_solver, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cint}, Ptr{Cdouble}),
                               ctxml, stream_ptr(), Cdouble(tol), Cint(itmx), n, r₂))
    push!(ml.n, n[]); nothing
end
function myproject!(a::Flow{n,T}, ml::CuMLPoisson{T}, w=1) where {n,T<:Union{Float32,Float64}}
    it = Ref{Cint}(0); r₂ = Ref{Cdouble}(0)
    ctxml = ml_handle(ml)
    check(context(ml.x), ccall((:ifadv_ml_myproject, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cint}, Ptr{Cdouble}),
                               ctxml, stream_ptr(), dptr(a.u), Cdouble(T(w) * last(a.Δt)), it, r₂))
    push!(ml.n, it[]); nothing
end

# ---- post-processing: level-set redistancing (src/redistaning.jl:31-87) and metric sums (src/metrics.jl) ---------------------------
function computeL!(L::CuArray{T}, ϕ::CuArray{T}, ϕini; perdir=()) where {T<:Union{Float32,Float64}}
    ctx = context(ϕ)
    check(ctx, ccall((:ifadv_redist_compute_l, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint),
                     ctx, stream_ptr(), dptr(L), dptr(ϕ), dptr(ϕini), perdir_mask(perdir))); nothing
end
function _redistaningStage!(ϕ::CuArray{T}, ϕ⁰, ϕini, L, dτ, α; perdir=()) where {T<:Union{Float32,Float64}}
    ctx = context(ϕ)
    check(ctx, ccall((:ifadv_redist_stage, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cuint),
                     ctx, stream_ptr(), dptr(ϕ), dptr(ϕ⁰), dptr(ϕini), dptr(L), dτ, α, perdir_mask(perdir))); nothing
end
function redistaning!(ls::LevelSet{D,T,<:CuArray}; d=5, dτ=0.5, perdir=()) where {D,T<:Union{Float32,Float64}}
    ctx = context(ls.ϕ)
    check(ctx, ccall((:ifadv_redistance, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cuint),
                     ctx, stream_ptr(), dptr(ls.ϕ), dptr(ls.ϕ⁰), dptr(ls.ϕini), dptr(ls.L), d, dτ, perdir_mask(perdir))); nothing
end
"""Σ over inside(f) of ρkeI, ρgh, ρuI(i) (src/metrics.jl:15-17,25,49-51) in one device pass: (ke, pe, momentum::NTuple{D})."""
function metric_sums(u::CuArray{T}, f::CuArray{T,D}, λρ; U=ntuple(_ -> 0.0, D), g=ntuple(_ -> 0.0, D), StatWL=ntuple(_ -> 0.0, D)) where {T<:Union{Float32,Float64},D}
    ctx = context(f)
    out = zeros(Cdouble, 5)
    pad(v) = Cdouble[v...; zeros(3 - D)]
    check(ctx, ccall((:ifadv_metrics, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                     ctx, stream_ptr(), dptr(u), dptr(f), λρ, pad(U), pad(g), pad(StatWL), out))
    out[1], out[2], ntuple(i -> out[2 + i], D)
end

# ---- MPFMomStep!  (src/flow.jl:60-109) with the transport half on the fused entry points ---------------------------------------
# Each of the two groups `copyto!(f⁰,f); u2ρu!(ρu,u⁰,·); BC!(ρu); advectfq!(…)` (flow.jl:61,69-70 and :89-92) becomes ONE call of
# ifadv_u2rhou_advect_vof_rhouu (bit-identical to the separate calls, include/ifadv.h); the forcing runs through the three overrides above, only udf! and the projection stay WaterLily's.
function fused_group!(a::Flow{D,T}, c::cVOF, fsrc, f, u¹, u², uOld, δt) where {D,T}
    a.uBC isa Function && error("IntfAdvB200Ext: function-valued uBC is not supported (no fallback)")
    ctx = context(f)
    dirv = Cint[ntuple(i -> mod(length(a.Δt) + i, D) + 1, D)...; zeros(Cint, 3 - D)]   # flow.jl:163
    A = Cdouble[a.uBC...; zeros(3 - D)]
    rep = Ref(IfadvReport())
    want = (CALLS[] += 1) % CHECK_EVERY[] == 0
    rc = ccall((:ifadv_u2rhou_advect_vof_rhouu, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid},
                Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint, Ptr{Cdouble}, Cuint, Cint, Ptr{Cint},
                Ptr{IfadvReport}),
               ctx, stream_ptr(), dptr(fsrc), dptr(f), dptr(c.fᶠ), dptr(a.σ), dptr(u¹), dptr(u²), δt, dptr(c.c̄),
               dptr(c.ρu), dptr(a.f), dptr(c.ρuf), dptr(uOld), dptr(c.dρ), c.λρ, limiter_enum(a.λ), normal_enum(c.normalScheme),
               A, perdir_mask(a.perdir), Cint(a.exitBC), dirv, want ? rep : C_NULL)
    report(check(ctx, rc), rep[])
    nothing
end
function InterfaceAdvection.MPFMomStep!(a::Flow{D,T}, b::WaterLily.AbstractPoisson, c::cVOF{D,T,<:CuArray}, d::WaterLily.AbstractBody;
                                        δt=last(a.Δt), udf=nothing, kwargs...) where {D,T<:Union{Float32,Float64}}
    t₁ = sum(a.Δt); t₀ = t₁ - δt; tₘ = t₁ - δt / 2
    stage!(fNow, tNow, tUdf, w, dtUdf) = begin   # forcing + projection of one RK stage (flow.jl:73-82 / :94-106)
        fill!(a.μ₀, 1)
        InterfaceAdvection.viscSurfTenρu!(a.f, a.u, a.σ, fNow, c.α, c.n̂, c.fᶠ, c.λμ, c.μ, c.λρ, c.η; perdir=a.perdir)
        u2ρu!(c.n̂, a.u⁰, c.f, c.λρ)
        w == 1 && (a.u⁰ .= a.u)
        w == 1 ? InterfaceAdvection.updateU!(a.u, c.ρu, c.n̂, a.f, δt, fNow, c.λρ, tNow, a.g, a.uBC) :
                 InterfaceAdvection.updateU!(a.u, c.ρu, c.n̂, a.f, δt, fNow, c.λρ, tNow, a.g, a.uBC, w)
        WaterLily.udf!(a, udf, a.u⁰, tUdf; dt=dtUdf, kwargs...)
        WaterLily.BC!(a.u, a.uBC, a.exitBC, a.perdir)
        InterfaceAdvection.updateL!(a.μ₀, fNow, c.λρ; perdir=a.perdir)
        WaterLily.update!(b)
        w == 1 ? InterfaceAdvection.myproject!(a, b) : InterfaceAdvection.myproject!(a, b, w)
        WaterLily.BC!(a.u, a.uBC, a.exitBC, a.perdir)
    end
    copyto!(a.u⁰, a.u)                                         # :61
    fused_group!(a, c, c.f, c.f⁰, a.u⁰, a.u, a.u, δt)          # :61 (f⁰←f), :69, :70
    @. c.f⁰ = (c.f⁰ + c.f) / 2                                 # :74
    stage!(c.f⁰, tₘ, t₀, T(1 / 2), T(1 / 2) * δt)              # :73-82
    copyto!(c.f⁰, c.f)                                         # :89
    fused_group!(a, c, c.f, c.f, a.u, a.u, a.u⁰, δt)           # :91, :92
    stage!(c.f, t₁, t₁, 1, δt)                                 # :94-106
    push!(a.Δt, min(MPCFL(a, c), 1.2δt))                      # :108
end

end # module
