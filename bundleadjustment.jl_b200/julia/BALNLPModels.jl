# BALNLPModels.jl -- drop-in replacement of src/BALNLPModels.jl that routes the hot path through libbagpu.so
# (include/bagpu.h).  Same type name, same constructor, same NLPModels methods as the reference file, so
# src/main.jl:5-30 and src/benchmark.jl run unchanged:
#
#     include("BALNLPModels.jl")          # this file instead of src/BALNLPModels.jl
#     include("lm.jl")                    # the reference's own Levenberg_Marquardt (src/lm.jl), untouched
#     include("lm_gpu.jl")                # adds the device method of Levenberg_Marquardt (optional)
#     BA    = BALNLPModel("LadyBug/problem-49-7776-pre.txt.bz2")
#     fr_BA = FeasibilityResidual(BA)
#     stats = Levenberg_Marquardt(fr_BA, :LDL, :Metis, :None, false)
#
# Written against NLPModels 0.12.4 (Manifest.toml:833-837), the version the reference pins.  NEVER EXECUTED in the
# build container (no Julia there): the Python mirror bundleadjustment.jl_b200/model.py binds the very same symbols
# with the same argument order and is what tests/ drives.
using NLPModels
include("ReadFiles.jl")                                   # src/ReadFiles.jl: readfile, unchanged

const libbagpu = get(ENV, "LIBBAGPU", "libbagpu.so")

struct BAError <: Exception
  code::Cint
  msg::String
end

# src/BALNLPModels.jl:79-88 plus the library handle
mutable struct BALNLPModel <: AbstractNLPModel
  meta::NLPModelMeta
  counters::Counters
  cams_indices::Vector{Int}
  pnts_indices::Vector{Int}
  pt2d::AbstractVector
  nobs::Int
  npnts::Int
  ncams::Int
  handle::Ptr{Cvoid}
end

function check(nlp::BALNLPModel, rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:ba_last_error, libbagpu), Cstring, (Ptr{Cvoid},), nlp.handle))
  throw(BAError(rc, msg))
end

# src/BALNLPModels.jl:58-68
function name(filename::AbstractString)
  k = 1
  while filename[k] != '/'
    k += 1
  end
  l = k + 8
  while filename[l] != 'p'
    l += 1
  end
  return filename[1:k-1] * filename[k+8:l-2]
end

# same constructor contract as src/BALNLPModels.jl:91-106.  ngpus = 1: one device (`device`); ngpus > 1 or :all: the
# library shards the observations over that many GPUs inside this one process (ba_create_multi) -- nothing else in
# the calling code changes.
function BALNLPModel(filename::AbstractString, T::Type = Float64; device::Integer = 0, ngpus = 1)
  T == Float64 || error("libbagpu computes in Float64 (BALNLPModel(filename, Float64))")
  cams_indices, pnts_indices, pt2d, x0, ncams, npnts, nobs = readfile(filename, T)   # src/ReadFiles.jl:9
  nvar, ncon = 9 * ncams + 3 * npnts, 2 * nobs
  meta = NLPModelMeta(nvar, ncon = ncon, x0 = x0, lcon = fill(0.0, ncon), ucon = fill(0.0, ncon),
                      nnzj = 2 * nobs * 12, name = name(filename))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  rc = if ngpus == 1
    ccall((:ba_create, libbagpu), Cint,
          (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
          ncams, npnts, nobs, cams_indices, pnts_indices, pt2d, device, h)
  else
    ccall((:ba_create_multi, libbagpu), Cint,
          (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}),
          ncams, npnts, nobs, cams_indices, pnts_indices, pt2d, ngpus === :all ? 0 : ngpus, C_NULL, h)
  end
  nlp = BALNLPModel(meta, Counters(), cams_indices, pnts_indices, pt2d, nobs, npnts, ncams, h[])
  h[] == C_NULL ? throw(BAError(rc, "ba_create failed")) : check(nlp, rc)
  finalizer(m -> ccall((:ba_destroy, libbagpu), Cint, (Ptr{Cvoid},), m.handle), nlp)
  @info "BALNLPModel $filename" nvar ncon
  return nlp
end

NLPModels.obj(::BALNLPModel, ::AbstractVector) = 0.0                                  # src/BALNLPModels.jl:109
NLPModels.grad!(::BALNLPModel, ::AbstractVector, g::AbstractVector) = fill!(g, 0)     # :112

# Page-locked vectors for the big outputs: the allocating forms cons(nlp, x) / jac_coord(nlp, x) that
# Levenberg_Marquardt calls once (src/lm.jl:39,54) hand back arrays backed by ba_alloc_pinned, so every later
# in-place call (src/lm.jl:252,268,341) copies at the full PCIe rate.  Ordinary arrays work too: the library
# then stages the copy through its own pinned ring (ba_hostio.cu).
function pinned_vector(::Type{T}, n::Integer) where {T}
  p = Ref{Ptr{Cvoid}}(C_NULL)
  ccall((:ba_alloc_pinned, libbagpu), Cint, (UInt64, Ref{Ptr{Cvoid}}), UInt64(n * sizeof(T)), p) == 0 ||
    error("ba_alloc_pinned failed")
  v = unsafe_wrap(Array, Ptr{T}(p[]), n; own = false)
  finalizer(a -> ccall((:ba_free_pinned, libbagpu), Cint, (Ptr{Cvoid},), pointer(a)), v)
  return v
end
NLPModels.cons(nlp::BALNLPModel, x::Vector{Float64}) = cons!(nlp, x, pinned_vector(Float64, nlp.meta.ncon))
NLPModels.jac_coord(nlp::BALNLPModel, x::Vector{Float64}) = jac_coord!(nlp, x, pinned_vector(Float64, nlp.meta.nnzj))

function NLPModels.cons!(nlp::BALNLPModel, x::Vector{Float64}, cx::Vector{Float64})           # :115-122
  increment!(nlp, :neval_cons)
  check(nlp, ccall((:ba_residual, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, cx))
  return cx
end

function NLPModels.jac_structure!(nlp::BALNLPModel, rows::Vector{Int}, cols::Vector{Int})      # :125-158
  increment!(nlp, :neval_jac)
  check(nlp, ccall((:ba_jac_structure, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nlp.handle, rows, cols))
  return rows, cols
end

function NLPModels.jac_coord!(nlp::BALNLPModel, x::Vector{Float64}, vals::Vector{Float64})      # :161-206
  increment!(nlp, :neval_jac)
  check(nlp, ccall((:ba_jac_coord, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, vals))
  return vals
end

# new on this surface (the reference forms these products with mul_sparse!, src/lma_aux.jl:194-212)
function NLPModels.jprod!(nlp::BALNLPModel, x::Vector{Float64}, v::Vector{Float64}, Jv::Vector{Float64})
  increment!(nlp, :neval_jprod)
  check(nlp, ccall((:ba_jprod, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   nlp.handle, x, v, Jv))
  return Jv
end

function NLPModels.jtprod!(nlp::BALNLPModel, x::Vector{Float64}, v::Vector{Float64}, Jtv::Vector{Float64})
  increment!(nlp, :neval_jtprod)
  check(nlp, ccall((:ba_jtprod, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   nlp.handle, x, v, Jtv))
  return Jtv
end

# knobs of the device solve (no counterpart in the reference; they never change the solution)
set_solver!(nlp::BALNLPModel, s::Symbol) =                                     # :auto, :pcg, :exact, :mixed
  check(nlp, ccall((:ba_set_solver, libbagpu), Cint, (Ptr{Cvoid}, Cint), nlp.handle,
                   Dict(:auto => 0, :pcg => 1, :exact => 2, :mixed => 3)[s]))
set_coarse_clusters!(nlp::BALNLPModel, n::Integer) =
  check(nlp, ccall((:ba_set_coarse_clusters, libbagpu), Cint, (Ptr{Cvoid}, Cint), nlp.handle, n))
set_deflation!(nlp::BALNLPModel, k::Integer) =
  check(nlp, ccall((:ba_set_deflation, libbagpu), Cint, (Ptr{Cvoid}, Cint), nlp.handle, k))
