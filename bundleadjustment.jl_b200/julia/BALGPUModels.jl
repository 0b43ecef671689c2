# BALGPUModels.jl -- the binding a maintainer of BundleAdjustment.jl adds to route the hot path through
# libbagpu.so (include/bagpu.h).  Written against NLPModels 0.12.4 (Manifest.toml:833-837), the version the
# reference pins.  NOT executed in the build container (no Julia there); the Python mirror in
# bundleadjustment.jl_b200/model.py binds the very same symbols and is what the tests drive.
#
# Usage, mirroring src/main.jl:8,27,30:
#     include("src/ReadFiles.jl"); include("BALGPUModels.jl")
#     BA    = BALGPUModel("LadyBug/problem-49-7776-pre.txt.bz2")
#     fr_BA = FeasibilityResidual(BA)                       # unchanged NLPModels adaptor
#     stats = Levenberg_Marquardt_GPU(fr_BA, :LDL, :AMD, :None, false)
using NLPModels, SolverTools

const libbagpu = get(ENV, "LIBBAGPU", "libbagpu.so")

struct BAError <: Exception
  code::Cint
  msg::String
end

mutable struct BALGPUModel <: AbstractNLPModel
  meta::NLPModelMeta
  counters::Counters
  cams_indices::Vector{Int}
  pnts_indices::Vector{Int}
  pt2d::Vector{Float64}
  nobs::Int
  npnts::Int
  ncams::Int
  handle::Ptr{Cvoid}
end

function check(nlp, rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:ba_last_error, libbagpu), Cstring, (Ptr{Cvoid},), nlp.handle))
  throw(BAError(rc, msg))
end

# same constructor contract as BALNLPModel(filename, T) -- src/BALNLPModels.jl:91-106
function BALGPUModel(filename::AbstractString; device::Integer = 0)
  cams_indices, pnts_indices, pt2d, x0, ncams, npnts, nobs = readfile(filename, Float64)   # src/ReadFiles.jl:9
  nvar, ncon = 9 * ncams + 3 * npnts, 2 * nobs
  meta = NLPModelMeta(nvar, ncon = ncon, x0 = x0, lcon = fill(0.0, ncon), ucon = fill(0.0, ncon),
                      nnzj = 2 * nobs * 12, name = name(filename))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:ba_create, libbagpu), Cint,
             (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
             ncams, npnts, nobs, cams_indices, pnts_indices, pt2d, device, h)
  nlp = BALGPUModel(meta, Counters(), cams_indices, pnts_indices, pt2d, nobs, npnts, ncams, h[])
  check(nlp, rc)
  finalizer(m -> ccall((:ba_destroy, libbagpu), Cint, (Ptr{Cvoid},), m.handle), nlp)
  return nlp
end

NLPModels.obj(::BALGPUModel, ::AbstractVector) = 0.0                       # src/BALNLPModels.jl:109
NLPModels.grad!(::BALGPUModel, ::AbstractVector, g::AbstractVector) = fill!(g, 0)   # :112

function NLPModels.cons!(nlp::BALGPUModel, x::Vector{Float64}, cx::Vector{Float64})           # :115-122
  increment!(nlp, :neval_cons)
  check(nlp, ccall((:ba_residual, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, cx))
  return cx
end

function NLPModels.jac_structure!(nlp::BALGPUModel, rows::Vector{Int}, cols::Vector{Int})    # :125-158
  increment!(nlp, :neval_jac)
  check(nlp, ccall((:ba_jac_structure, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nlp.handle, rows, cols))
  return rows, cols
end

function NLPModels.jac_coord!(nlp::BALGPUModel, x::Vector{Float64}, vals::Vector{Float64})    # :161-206
  increment!(nlp, :neval_jac)
  check(nlp, ccall((:ba_jac_coord, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.handle, x, vals))
  return vals
end

# new on this surface (the reference forms these products with mul_sparse!, src/lma_aux.jl:194-212)
function NLPModels.jprod!(nlp::BALGPUModel, x::Vector{Float64}, v::Vector{Float64}, Jv::Vector{Float64})
  increment!(nlp, :neval_jprod)
  check(nlp, ccall((:ba_jprod, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   nlp.handle, x, v, Jv))
  return Jv
end

function NLPModels.jtprod!(nlp::BALGPUModel, x::Vector{Float64}, v::Vector{Float64}, Jtv::Vector{Float64})
  increment!(nlp, :neval_jtprod)
  check(nlp, ccall((:ba_jtprod, libbagpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   nlp.handle, x, v, Jtv))
  return Jtv
end

# PCG preconditioner knobs of the device solve (no counterpart in the reference; they change iteration counts only)
set_coarse_clusters!(nlp::BALGPUModel, n::Integer) =
  check(nlp, ccall((:ba_set_coarse_clusters, libbagpu), Cint, (Ptr{Cvoid}, Cint), nlp.handle, n))
set_deflation!(nlp::BALGPUModel, k::Integer) =
  check(nlp, ccall((:ba_set_deflation, libbagpu), Cint, (Ptr{Cvoid}, Cint), nlp.handle, k))

# ba_lm_params / ba_lm_stats / ba_lm_row of include/bagpu.h (isbits structs, same field order)
struct BALMParams
  restol::Float64; satol::Float64; srtol::Float64; oatol::Float64; ortol::Float64; atol::Float64; rtol::Float64
  nu_d::Float64; nu_m::Float64; lambda::Float64; delta_d::Float64
  ite_max::Int64; linesearch::Int32; pcg_max_iter::Int32; pcg_tol::Float64
end
struct BALMStats
  status::Int32; pad::Int32; iter::Int64
  objective::Float64; dual_feas::Float64; lambda_final::Float64; elapsed_s::Float64
  pcg_iters_total::Int64
  t_eval_ms::Float64; t_assemble_ms::Float64; t_pcg_ms::Float64; t_backsub_ms::Float64
end
struct BALMRow
  iter::Int64; f::Float64; df::Float64; dfeas::Float64; lambda::Float64; delta_norm::Float64; rho::Float64
  accepted::Int32; acc_str::Int32; pcg_iters::Int32; ntimes::Int32
end

const LM_STATUS = (:unknown, :small_step, :first_order, :small_residual, :acceptable, :neg_pred, :exception, :max_iter)

function lm_log_row(rowp::Ptr{BALMRow}, ::Ptr{Cvoid})::Cvoid     # same 8 columns as src/lm.jl:304
  r = unsafe_load(rowp)
  @info log_row(Any[r.iter, r.f, r.df, r.dfeas, r.lambda, r.delta_norm, r.rho, r.acc_str != 0 ? "acc" : "rej"])
  return
end

# Same signature and defaults as Levenberg_Marquardt (src/lm.jl:15-26).  facto / perm / normalize are accepted
# and ignored: each combination solves (J'J + λI) δ = -J'r, which the library solves on the GPU.
function Levenberg_Marquardt_GPU(model::AbstractNLSModel, facto::Symbol, perm::Symbol, normalize::Symbol,
                                 linesearch::Bool; x::Vector{Float64} = copy(model.meta.x0),
                                 restol = eps(Float64)^(1/3), satol = sqrt(eps(Float64)), srtol = sqrt(eps(Float64)),
                                 oatol = sqrt(eps(Float64)), ortol = eps(Float64)^(1/3),
                                 atol = sqrt(eps(Float64)), rtol = eps(Float64)^(1/3),
                                 νd = 3.0, νm = 3.0, λ = 30.0, δd = 2.0, ite_max::Int = 200, max_time::Int = 3600,
                                 pcg_tol = 1e-13, pcg_max_iter = 1000)
  nlp = model.nlp::BALGPUModel                      # FeasibilityResidual keeps the wrapped model in .nlp
  prm = BALMParams(restol, satol, srtol, oatol, ortol, atol, rtol, νd, νm, λ, δd, ite_max, linesearch, pcg_max_iter, pcg_tol)
  st = Ref{BALMStats}()
  cb = @cfunction(lm_log_row, Cvoid, (Ptr{BALMRow}, Ptr{Cvoid}))
  @info log_header([:iter, :f, :df, :dfeas, :λ, :δ, :ρ, :status], [Int, Float64, Float64, Float64, Float64, Float64, Float64, String])
  t0 = time()
  check(nlp, ccall((:ba_lm_solve, libbagpu), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ref{BALMParams}, Ref{BALMStats}, Ptr{Cvoid}, Ptr{Cvoid}),
                   nlp.handle, x, prm, st, cb, C_NULL))
  s = st[]
  return GenericExecutionStats(LM_STATUS[s.status + 1], model, solution = x, objective = s.objective,
                               iter = Int(s.iter), elapsed_time = time() - t0, dual_feas = s.dual_feas)   # src/lm.jl:409-415
end
