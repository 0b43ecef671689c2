# lm_gpu.jl -- the device METHOD of the reference's own function `Levenberg_Marquardt` (src/lm.jl:15-26).
#
# include it after src/lm.jl.  It is a method of the same generic function with a more specific first argument
# (FeasibilityResidual is a concrete subtype of AbstractNLSModel, NLPModels 0.12.4), so the call in src/main.jl:30,
#     stats = Levenberg_Marquardt(fr_BA, :LDL, :Metis, :None, false)
# reaches it without any edit.  If the wrapped model is not a libbagpu BALNLPModel it forwards to the reference's
# generic method with `invoke`; otherwise the whole loop (src/lm.jl:102-405, decision for decision) runs on the
# device inside ba_lm_solve.  facto / perm / normalize are accepted as they are: every combination of the reference
# solves (J'J + λI) δ = -J'r (SURVEY.md section 3.4), which the library solves exactly (explicit reduced camera
# system + dense Cholesky, up to 2048 cameras) or by matrix-free PCG.
# NEVER EXECUTED in the build container (no Julia); bundleadjustment.jl_b200/lm.py is the tested mirror.
using NLPModels, SolverTools

# ba_lm_params / ba_lm_stats / ba_lm_row of include/bagpu.h (isbits structs, same field order;
# tests/test_host.py::test_struct_layouts_match_the_header checks the Python mirrors of the same layout)
struct BALMParams
  restol::Float64; satol::Float64; srtol::Float64; oatol::Float64; ortol::Float64; atol::Float64; rtol::Float64
  nu_d::Float64; nu_m::Float64; lambda::Float64; delta_d::Float64
  ite_max::Int64; linesearch::Int32; pcg_max_iter::Int32; pcg_tol::Float64
  solver::Int32; reserved::Int32
end
struct BALMStats
  status::Int32; pad::Int32; iter::Int64
  objective::Float64; dual_feas::Float64; lambda_final::Float64; elapsed_s::Float64
  pcg_iters_total::Int64
  t_eval_ms::Float64; t_assemble_ms::Float64; t_pcg_ms::Float64; t_backsub_ms::Float64
  capped_solves::Int64; worst_solve_rel::Float64; t_prepare_ms::Float64
  t_schur_ms::Float64; t_chol_ms::Float64; chol_n::Int64; chol_count::Int64
  mixed_fallbacks::Int64
end
struct BALMRow
  iter::Int64; f::Float64; df::Float64; dfeas::Float64; lambda::Float64; delta_norm::Float64; rho::Float64
  accepted::Int32; acc_str::Int32; pcg_iters::Int32; ntimes::Int32
  solver::Int32; converged::Int32; solve_rel::Float64
end

const LM_STATUS = (:unknown, :small_step, :first_order, :small_residual, :acceptable, :neg_pred, :exception, :max_iter)

function lm_log_row(rowp::Ptr{BALMRow}, ::Ptr{Cvoid})::Cvoid     # the same 8 columns as src/lm.jl:304
  r = unsafe_load(rowp)
  @info log_row(Any[r.iter, r.f, r.df, r.dfeas, r.lambda, r.delta_norm, r.rho, r.acc_str != 0 ? "acc" : "rej"])
  r.converged == 0 && @warn "damped solve stopped at pcg_max_iter (relative residual $(r.solve_rel)): inexact step"
  return
end

function Levenberg_Marquardt(model::FeasibilityResidual, facto::Symbol, perm::Symbol, normalize::Symbol,
                             linesearch::Bool; x::AbstractVector = copy(model.meta.x0), facto_type::DataType = eltype(x),
                             restol = eps(Float64)^(1/3), satol = sqrt(eps(Float64)), srtol = sqrt(eps(Float64)),
                             oatol = sqrt(eps(Float64)), ortol = eps(Float64)^(1/3),
                             atol = sqrt(eps(Float64)), rtol = eps(Float64)^(1/3),
                             νd = 3.0, νm = 3.0, λ = 30.0, δd = 2.0, ite_max::Int = 200, max_time::Int = 3600,
                             solver::Symbol = :auto, pcg_tol = 1e-13, pcg_max_iter = 1000)
  if !(model.nlp isa BALNLPModel && hasfield(typeof(model.nlp), :handle))
    # not a libbagpu model: the reference's own method (src/lm.jl:15)
    return invoke(Levenberg_Marquardt, Tuple{AbstractNLSModel, Symbol, Symbol, Symbol, Bool}, model, facto, perm,
                  normalize, linesearch; x = x, facto_type = facto_type, restol = restol, satol = satol, srtol = srtol,
                  oatol = oatol, ortol = ortol, atol = atol, rtol = rtol, νd = νd, νm = νm, λ = λ, δd = δd,
                  ite_max = ite_max, max_time = max_time)
  end
  nlp = model.nlp::BALNLPModel
  xs = Vector{Float64}(x)                                                 # x0 in, solution out
  prm = BALMParams(restol, satol, srtol, oatol, ortol, atol, rtol, νd, νm, λ, δd, ite_max, linesearch,
                   pcg_max_iter, pcg_tol, Dict(:auto => 0, :pcg => 1, :exact => 2, :mixed => 3)[solver], 0)
  st = Ref{BALMStats}()
  cb = @cfunction(lm_log_row, Cvoid, (Ptr{BALMRow}, Ptr{Cvoid}))
  @info log_header([:iter, :f, :df, :dfeas, :λ, :δ, :ρ, :status], [Int, Float64, Float64, Float64, Float64, Float64, Float64, String])
  t0 = time()
  check(nlp, ccall((:ba_lm_solve, libbagpu), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ref{BALMParams}, Ref{BALMStats}, Ptr{Cvoid}, Ptr{Cvoid}),
                   nlp.handle, xs, prm, st, cb, C_NULL))
  s = st[]
  return GenericExecutionStats(LM_STATUS[s.status + 1], model, solution = xs, objective = s.objective,
                               iter = Int(s.iter), elapsed_time = time() - t0, dual_feas = s.dual_feas)   # src/lm.jl:409-415
end
