# src/cuda.jl — the Julia side of the B200 hot path: thin `ccall` bindings to libfastcheb_cuda
# (C ABI: include/fastcheb.h).  Drop this file into FastChebInterp.jl's src/ and add
# `include("cuda.jl")` after `include("eval.jl")` in src/FastChebInterp.jl (line 62).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  Every entry point
# used here is exercised through the identical C ABI from Python (fastchebinterp.jl_b200/_capi.py,
# tests/test_gpu_*.py).  Memory layouts are those of SURVEY.md Appendix A and were chosen so that no
# copy or permutation is needed on the Julia side: `Array{SVector{K,T},N}` coefficients,
# `Vector{SVector{N,T}}` points, `Vector{Tuple{V,SMatrix{K,N}}}` Jacobian results are passed as is.
#
# Public API (unchanged names; METHODS are added to the reference's own functions for the device-resident types):
#   cu = CuChebPoly(c::ChebPoly)            upload (FastChebInterp.jl:37-41 -> fastcheb_create)
#   cu = CuChebPoly(c; devices=0:7)         replicated on several GPUs, driven from this one process (fastcheb_multi_*)
#   cu(ys::AbstractVector{<:SVector{N}})    == c.(ys)                      (eval.jl:63-70); mixed precisions promote (eval.jl:25)
#   cu.(ys)                                 same, through Base.broadcasted
#   chebjacobian(cu, ys), chebgradient(cu, ys)                             (eval.jl:147-177)
#   chebinterp(ondevice(vals), lb, ub; tol) -> CuChebPoly                  (interp.jl:95-99,140-141): a method of chebinterp
#   chebinterp(cu, order, lb, ub; tol)      re-interpolation on the device (interp.jl:163-170; the loop body of roots)
#   chebcoefs_batched(vals::Matrix)         10^5 independent 1-D fits      (interp.jl:39-57 per column)
#   ChebPoly(cu)                            download

using StaticArrays

const libfastcheb = get(ENV, "FASTCHEB_CUDA_LIB", "libfastcheb_cuda")

const FASTCHEB_F32, FASTCHEB_F64 = Cint(0), Cint(1)
const FASTCHEB_MEM_HOST, FASTCHEB_MEM_DEVICE = Cint(0), Cint(1)
const FASTCHEB_OK, FASTCHEB_EDOMAIN, FASTCHEB_ENONFINITE, FASTCHEB_EDIMMISMATCH, FASTCHEB_EINVAL =
    Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)

_dtype(::Type{Float32}) = FASTCHEB_F32
_dtype(::Type{Float64}) = FASTCHEB_F64
# (real scalar type, is_complex, components) of a coefficient element type
_elt(::Type{T}) where {T<:Union{Float32,Float64}} = (T, false, 1)
_elt(::Type{Complex{T}}) where {T<:Union{Float32,Float64}} = (T, true, 1)
_elt(::Type{SVector{K,T}}) where {K,T} = (_elt(T)[1], _elt(T)[2], K)

function _last_error()
    buf = Vector{UInt8}(undef, 512)
    n = ccall((:fastcheb_last_error, libfastcheb), Csize_t, (Ptr{UInt8}, Csize_t), buf, length(buf))
    String(buf[1:min(n, length(buf) - 1)])
end
function _error_index()
    idx = Ref{Int64}(0)
    ccall((:fastcheb_error_index, libfastcheb), Cint, (Ref{Int64},), idx)
    idx[]
end

# status code -> the exception the reference throws at the same place
function _check(rc::Cint, c=nothing, xs=nothing)
    rc == FASTCHEB_OK && return
    if rc == FASTCHEB_EDOMAIN                      # eval.jl:64, :148
        i = _error_index() + 1
        x = xs === nothing ? "point $i" : xs[i]
        throw(ArgumentError("$x not in domain" * (c === nothing ? "" : " $(c.lb) to $(c.ub)")))
    elseif rc == FASTCHEB_ENONFINITE               # interp.jl:40
        throw(DomainError(_error_index() + 1, "non-finite interpolant value"))
    elseif rc == FASTCHEB_EDIMMISMATCH             # eval.jl:170
        throw(DimensionMismatch(_last_error()))
    elseif rc == FASTCHEB_EINVAL
        throw(ArgumentError(_last_error()))
    else
        error("libfastcheb_cuda: ", _last_error(), " (status $rc)")
    end
end

"""
Device-resident `ChebPoly{N,T,Td}`: opaque handles owned by libfastcheb_cuda (coefficients in the kernels' own layouts,
`lb`, `ub`, sizes).  Immutable after creation; safe to evaluate from several tasks at once.  `multi != C_NULL`: the
polynomial is replicated on several GPUs (`CuChebPoly(c; devices=0:7)`) and every batch is block-partitioned over them
from this one process.  A `Float32` polynomial evaluated at `Float64` points promotes like the reference
(`muladd(x, c, c)`, eval.jl:25): the widened `Float64` handle is created on first use and cached in `wide`.
"""
mutable struct CuChebPoly{N,T,Td<:Real} <: Function
    handle::Ptr{Cvoid}      # fastcheb_poly (single GPU) or C_NULL
    multi::Ptr{Cvoid}       # fastcheb_multi or C_NULL
    lb::SVector{N,Td}
    ub::SVector{N,Td}
    dims::NTuple{N,Int}
    devices::Vector{Int}
    wide::Any               # nothing, or the CuChebPoly of the Float64-widened coefficients
    function CuChebPoly{N,T,Td}(h, m, lb, ub, dims, devices) where {N,T,Td}
        c = new{N,T,Td}(h, m, lb, ub, dims, devices, nothing)
        finalizer(c) do x   # destroy entry points are finalizer-safe: no exceptions, any thread, NULL is a no-op
            ccall((:fastcheb_destroy, libfastcheb), Cint, (Ptr{Cvoid},), x.handle)
            ccall((:fastcheb_multi_destroy, libfastcheb), Cint, (Ptr{Cvoid},), x.multi)
        end
        c
    end
end
Base.ndims(::CuChebPoly{N}) where {N} = N

function CuChebPoly(c::ChebPoly{N,T,Td}; device::Integer=0, devices=nothing) where {N,T,Td}
    R, cplx, K = _elt(T)
    dims = collect(Int64, size(c.coefs))
    lb, ub = collect(Float64, c.lb), collect(Float64, c.ub)   # Td may be Int (runtests.jl:170)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if devices === nothing
        GC.@preserve c begin
            rc = ccall((:fastcheb_create, libfastcheb), Cint,
                       (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Cint),
                       h, N, dims, _dtype(R), cplx, K, pointer(c.coefs), FASTCHEB_MEM_HOST, lb, ub, device)
        end
        _check(rc)
        return CuChebPoly{N,T,Td}(h[], C_NULL, c.lb, c.ub, size(c.coefs), [Int(device)])
    end
    devs = collect(Cint, devices)
    GC.@preserve c begin
        rc = ccall((:fastcheb_multi_create, libfastcheb), Cint,
                   (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}, Cint, Ptr{Int64}, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}),
                   h, length(devs), devs, N, dims, _dtype(R), cplx, K, pointer(c.coefs), FASTCHEB_MEM_HOST, lb, ub)
    end
    _check(rc)
    CuChebPoly{N,T,Td}(C_NULL, h[], c.lb, c.ub, size(c.coefs), collect(Int, devices))
end

function ChebPoly(cu::CuChebPoly{N,T,Td}) where {N,T,Td}
    coefs = Array{T,N}(undef, cu.dims)
    h = cu.handle != C_NULL ? cu.handle : ccall((:fastcheb_multi_poly, libfastcheb), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), cu.multi, 0)
    _check(ccall((:fastcheb_get_coefs, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), h, coefs, FASTCHEB_MEM_HOST))
    ChebPoly{N,T,Td}(coefs, cu.lb, cu.ub)
end

# result element types — the same promotion as eval.jl:25 (`muladd(x, c, c)`)
_valtype(::Type{T}, ::Type{Tx}) where {T,Tx} = typeof(zero(T) * zero(Tx))
_widen(::Type{Float32}) = Float64
_widen(::Type{Complex{Float32}}) = Complex{Float64}
_widen(::Type{SVector{K,T}}) where {K,T} = SVector{K,_widen(T)}

# the handle that computes in the precision of the points: Float32 coefficients at Float64 points promote (eval.jl:25)
function _for_points(cu::CuChebPoly{N,T,Td}, ::Type{R}) where {N,T,Td,R}
    _elt(T)[1] === R && return cu
    if _elt(T)[1] === Float32 && R === Float64
        if cu.wide === nothing
            c = ChebPoly(cu)
            cw = ChebPoly{N,_widen(T),Td}(map(x -> _widen(T)(x), c.coefs), c.lb, c.ub)
            cu.wide = length(cu.devices) > 1 || cu.multi != C_NULL ? CuChebPoly(cw; devices=cu.devices) : CuChebPoly(cw; device=cu.devices[1])
        end
        return cu.wide
    end
    nothing   # Float64 coefficients at Float32 points: the points are widened by the caller below
end

# ---- c.(ys): batched evaluation (eval.jl:63-70 per point)
function (cu::CuChebPoly{N,T})(ys::AbstractVector{SVector{N,R}}) where {N,T,R<:Union{Float32,Float64}}
    h = _for_points(cu, R)
    h === nothing && return cu(map(y -> Float64.(y), ys))        # promote the points instead (eval.jl:65)
    h === cu || return h(ys)
    ys = ys isa Vector ? ys : collect(ys)
    out = Vector{T}(undef, length(ys))
    GC.@preserve ys out begin
        rc = cu.multi != C_NULL ?
            ccall((:fastcheb_multi_eval, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint),
                  cu.multi, pointer(ys), length(ys), pointer(out), FASTCHEB_MEM_HOST) :
            ccall((:fastcheb_eval, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                  cu.handle, pointer(ys), length(ys), pointer(out), FASTCHEB_MEM_HOST, C_NULL)
    end
    _check(rc, cu, ys)
    out
end
(cu::CuChebPoly{1,T})(xs::AbstractVector{R}) where {T,R<:Union{Float32,Float64}} =
    cu(reinterpret(SVector{1,R}, xs isa Vector ? xs : collect(xs)))
(cu::CuChebPoly{N})(ys::AbstractVector{<:SVector{N,<:Real}}) where {N} = cu(map(y -> float.(y), ys))   # integer points etc.
(cu::CuChebPoly{N})(x::SVector{N,<:Real}) where {N} = cu([float.(x)])[1]       # single point: still the device path
(cu::CuChebPoly{1})(x::Real) = cu(SVector(float(x)))
# `cu.(ys)` lowers to broadcasted(cu, ys): route whole batches to one library call
Base.Broadcast.broadcasted(cu::CuChebPoly{N}, ys::AbstractVector{<:SVector{N}}) where {N} = cu(ys)
Base.Broadcast.broadcasted(cu::CuChebPoly{1}, xs::AbstractVector{<:Real}) = cu(xs)

# ---- chebjacobian / chebgradient over batches (eval.jl:147-177).  The library writes Julia's own
# Vector{Tuple{V,SMatrix{K,N}}} element layout (fastcheb_eval_jac_tuple), so the result needs no repacking.
function chebjacobian(cu::CuChebPoly{N,T}, ys::AbstractVector{SVector{N,R}}) where {N,T,R<:Union{Float32,Float64}}
    h = _for_points(cu, R)
    h === nothing && return chebjacobian(cu, map(y -> Float64.(y), ys))
    h === cu || return chebjacobian(h, ys)
    _, _, K = _elt(T)
    S = T <: SVector ? eltype(T) : T
    out = Vector{Tuple{T,SMatrix{K,N,S,K * N}}}(undef, length(ys))
    ys = ys isa Vector ? ys : collect(ys)
    GC.@preserve ys out begin
        rc = _jac_tuple(cu, ys, out)
    end
    _check(rc, cu, ys)
    out
end
_jac_tuple(cu, ys, out) = cu.multi != C_NULL ?
    ccall((:fastcheb_multi_eval_jac_tuple, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint),
          cu.multi, pointer(ys), length(ys), pointer(out), FASTCHEB_MEM_HOST) :
    ccall((:fastcheb_eval_jac_tuple, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
          cu.handle, pointer(ys), length(ys), pointer(out), FASTCHEB_MEM_HOST, C_NULL)
function chebgradient(cu::CuChebPoly{N,T}, ys::AbstractVector{SVector{N,R}}) where {N,T,R<:Union{Float32,Float64}}
    T <: SVector && length(T) != 1 && throw(DimensionMismatch())                 # eval.jl:170
    h = _for_points(cu, R)
    h === nothing && return chebgradient(cu, map(y -> Float64.(y), ys))
    h === cu || return chebgradient(h, ys)
    out = Vector{Tuple{T,SVector{N,T}}}(undef, length(ys))                       # same bytes as (v, 1xN SMatrix)
    ys = ys isa Vector ? ys : collect(ys)
    GC.@preserve ys out begin
        rc = _jac_tuple(cu, ys, out)
    end
    _check(rc, cu, ys)
    out
end
chebjacobian(cu::CuChebPoly{N}, x::SVector{N,<:Real}) where {N} = chebjacobian(cu, [float.(x)])[1]
chebgradient(cu::CuChebPoly{N}, x::SVector{N,<:Real}) where {N} = chebgradient(cu, [float.(x)])[1]
chebgradient(cu::CuChebPoly{1}, x::Real) = (r = chebgradient(cu, SVector(float(x))); (r[1], r[2][1]))  # eval.jl:174-177

# ---- chebinterp(vals, lb, ub; tol) on the device (interp.jl:95-99, 140-141): fit + droptol + handle.
# The reference's method chebinterp(vals::AbstractArray, lb, ub; tol) stays what it is (CPU / FFTW); the device path is
# a METHOD OF THE SAME FUNCTION selected by the argument type: wrap the values with `ondevice(vals; device=0)`.
struct OnDevice{A<:AbstractArray}
    data::A
    device::Int
end
ondevice(vals::AbstractArray; device::Integer=0) = OnDevice(vals, Int(device))

function chebinterp(v::OnDevice{<:AbstractArray{T,N}}, lb, ub; tol::Real=-1) where {T,N}
    vals, device = v.data, v.device
    R, cplx, K = _elt(T)
    Td = promote_type(eltype(lb), eltype(ub))
    vals = vals isa Array ? vals : collect(vals)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dims = collect(Int64, size(vals))
    GC.@preserve vals begin
        rc = ccall((:fastcheb_interp, libfastcheb), Cint,
                   (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Cdouble, Cint, Ptr{Cvoid}),
                   h, N, dims, _dtype(R), cplx, K, pointer(vals), FASTCHEB_MEM_HOST,
                   collect(Float64, lb), collect(Float64, ub), tol, device, C_NULL)    # tol < 0: eps(R), interp.jl:102
    end
    _check(rc)
    nd = Vector{Int64}(undef, N)
    _check(ccall((:fastcheb_dims, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Int64}), h[], nd))
    CuChebPoly{N,T,Td}(h[], C_NULL, SVector{N,Td}(lb), SVector{N,Td}(ub), Tuple(nd), [device])
end
chebinterp(v::OnDevice{<:AbstractVector}, lb::Real, ub::Real; kws...) = chebinterp(v, SVector(lb), SVector(ub); kws...)   # interp.jl:140-141

# ---- many independent 1-D fits at once (BASELINE configs[4]): column j of `vals` is one line
function chebcoefs_batched(vals::Matrix{R}; device::Integer=0, devices=nothing) where {R<:Union{Float32,Float64}}
    n, howmany = size(vals)
    coefs = similar(vals)
    maxabs = Vector{Float64}(undef, howmany)
    if devices !== nothing        # the batch split per GPU, one process (SURVEY.md 8e: fits are sharded, a single fit is not)
        devs = collect(Cint, devices)
        GC.@preserve vals coefs maxabs begin
            rc = ccall((:fastcheb_multi_fit, libfastcheb), Cint,
                       (Cint, Ptr{Cint}, Cint, Ptr{Int64}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint),
                       length(devs), devs, 1, Int64[n], _dtype(R), 0, 1, howmany, pointer(vals), pointer(coefs), pointer(maxabs),
                       FASTCHEB_MEM_HOST)
        end
        _check(rc)
        return coefs, maxabs
    end
    GC.@preserve vals coefs maxabs begin
        rc = ccall((:fastcheb_fit, libfastcheb), Cint,
                   (Cint, Ptr{Int64}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Cvoid}),
                   1, Int64[n], _dtype(R), 0, 1, howmany, pointer(vals), pointer(coefs), pointer(maxabs),
                   FASTCHEB_MEM_HOST, device, C_NULL)
    end
    _check(rc)
    coefs, maxabs      # droptol per column: trailing |c| < tol*maxabs[j] (interp.jl:77-93)
end

# ---- callers of the hot path (SURVEY.md 8f): the same pattern, one ccall each ------------------------------------

# chebinterp(c, order, lb, ub; tol) with a device-resident ChebPoly as the function (interp.jl:163-170; the loop body of
# roots, roots.jl:105-107): grid -> evaluation -> fit -> droptol never leave the GPU.
# A method of the reference's chebinterp(f, order, lb, ub): CuChebPoly <: Function, and this more specific method wins.
function chebinterp(cu::CuChebPoly{N,T,Td}, order::NTuple{N,Int}, lb, ub; tol::Real=-1) where {N,T,Td}
    cu.handle == C_NULL && error("re-interpolation runs on one GPU: build the source with CuChebPoly(c; device=d)")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:fastcheb_reinterp, libfastcheb), Cint,
                 (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Cvoid}),
                 h, cu.handle, collect(Int64, order), collect(Float64, lb), collect(Float64, ub), tol, C_NULL))
    nd = Vector{Int64}(undef, N)
    _check(ccall((:fastcheb_dims, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Int64}), h[], nd))
    CuChebPoly{N,T,Td}(h[], C_NULL, SVector{N,Td}(lb), SVector{N,Td}(ub), Tuple(nd), cu.devices)
end

# rrule / frule (chainrules.jl:4-39) over batches: the Jacobian is contracted inside the kernel.
function cheb_pullback(cu::CuChebPoly{N,T}, ys::Vector{SVector{N,R}}, dy::Vector{T}) where {N,T,R<:Union{Float32,Float64}}
    v = Vector{T}(undef, length(ys)); xbar = Vector{SVector{N,R}}(undef, length(ys))
    GC.@preserve ys dy v xbar begin
        rc = ccall((:fastcheb_eval_vjp, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                   cu.handle, pointer(ys), length(ys), pointer(dy), pointer(v), pointer(xbar), FASTCHEB_MEM_HOST, C_NULL)
    end
    _check(rc, cu, ys)
    v, xbar            # real(J' * dy) per point (chainrules.jl:7,14)
end
function cheb_pushforward(cu::CuChebPoly{N,T}, ys::Vector{SVector{N,R}}, dx::Vector{SVector{N,R}}) where {N,T,R<:Union{Float32,Float64}}
    v = Vector{T}(undef, length(ys)); ydot = Vector{T}(undef, length(ys))
    GC.@preserve ys dx v ydot begin
        rc = ccall((:fastcheb_eval_jvp, libfastcheb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                   cu.handle, pointer(ys), length(ys), pointer(dx), pointer(v), pointer(ydot), FASTCHEB_MEM_HOST, C_NULL)
    end
    _check(rc, cu, ys)
    v, ydot            # J * dx per point (chainrules.jl:24)
end

# chebvandermonde (regression.jl:5-46) assembled on the device in one O(m P) pass; `A \ Y` (regression.jl:67) stays Julia.
function cuchebvandermonde(x::Vector{SVector{N,R}}, lb::SVector{N}, ub::SVector{N}, order::NTuple{N,Int}) where {N,R<:Union{Float32,Float64}}
    A = Matrix{R}(undef, length(x), prod(order .+ 1))
    GC.@preserve x A begin
        rc = ccall((:fastcheb_vandermonde, libfastcheb), Cint,
                   (Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Cvoid}),
                   N, collect(Int64, order .+ 1), _dtype(R), pointer(x), length(x), collect(Float64, lb), collect(Float64, ub),
                   pointer(A), size(A, 1), FASTCHEB_MEM_HOST, 0, C_NULL)
    end
    rc == FASTCHEB_EDOMAIN && throw(ArgumentError("$(x[_error_index() + 1]) not in domain [$lb,$ub]"))   # regression.jl:33
    _check(rc)
    A
end

# many independent 1-D polynomials, each at its own points: [c.(xs) for (c, xs) in zip(cs, xss)] in one launch.
# coefs: n x howmany (column j = polynomial j, e.g. the output of chebcoefs_batched), xs: M x howmany.
function cheb_eval_many(coefs::Matrix{R}, lb::Real, ub::Real, xs::Matrix{R}) where {R<:Union{Float32,Float64}}
    n, howmany = size(coefs); M = size(xs, 1)
    out = similar(xs)
    GC.@preserve coefs xs out begin
        rc = ccall((:fastcheb_eval_many, libfastcheb), Cint,
                   (Int64, Int64, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
                   howmany, n, _dtype(R), 0, 1, pointer(coefs), Float64[lb], Float64[ub], 0, pointer(xs), M, pointer(out), C_NULL, C_NULL,
                   FASTCHEB_MEM_HOST, 0, C_NULL)
    end
    _check(rc)
    out
end

# ... and at ONE shared set of points: [c.(xs) for c in cs] is then a matrix product on the FP64 tensor cores
# (fastcheb_eval_many_shared; n <= 65).  Result: M x howmany.
function cheb_eval_many(coefs::Matrix{R}, lb::Real, ub::Real, xs::Vector{R}) where {R<:Union{Float32,Float64}}
    n, howmany = size(coefs); M = length(xs)
    out = Matrix{R}(undef, M, howmany)
    GC.@preserve coefs xs out begin
        rc = ccall((:fastcheb_eval_many_shared, libfastcheb), Cint,
                   (Int64, Int64, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
                   howmany, n, _dtype(R), 0, 1, pointer(coefs), Float64[lb], Float64[ub], pointer(xs), M, pointer(out), C_NULL, C_NULL,
                   FASTCHEB_MEM_HOST, 0, C_NULL)
    end
    rc == FASTCHEB_EDOMAIN && throw(ArgumentError("$(xs[_error_index() + 1]) not in domain $lb to $ub"))
    _check(rc)
    out
end
