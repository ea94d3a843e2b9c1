# baseline/julia_ref.jl -- pins parity against the REAL reference.
#
# Run with a Julia >= 1.0 that has NLsolve, ForwardDiff and Dierckx installed (the only third-party packages the
# discrete-vortex path needs; plotting / UI packages are NOT loaded):
#
#     julia baseline/julia_ref.jl /path/to/UnsteadyFlowSolvers.jl tests/golden/julia
#
# It builds a minimal module from the reference's own, unmodified hot-path sources
# (src/kinem.jl, utils.jl, delVort.jl, lowOrder2D/{typedefs,calcs,solvers,postprocess}.jl), after two text patches
# without which the snapshot does not load / run (SURVEY.md section 4):
#   * src/lowOrder2D/solvers.jl:451   `if startflag == 0ml`  ->  `if startflag == 0`
#   * every `nlsolve(not_in_place(f), x0)` call is routed through `traced_nlsolve` below, which calls the very same
#     NLsolve.nlsolve but records every point the functor is evaluated at (solvers.jl:91,199,216,718,737)
# (src/plots/plots2D.jl, whose missing `end` breaks `using UnsteadyFlowSolvers`, is simply not included.)
#
# Outputs (JSON, one file per case) into the target directory; tests/test_julia_golden.py loads them, checks the CPU
# oracle in BOTH shedding-solve modes and the GPU against them, and reports which mode matches.  Nothing here is
# executed in this repository's image (there is no Julia); the file is kept so that anyone with the reference's
# toolchain can pin the oracle in one command.
#
# Cases (SURVEY.md 8d): C1 = test/ldvmLinTest.jl through ldvm / ldvmLin / lautat (468 steps);
# C1r-150 = Eldredge ramp to 45 deg, K = 0.2, pivot at the LE, LESPcrit 0.11, 150 steps (LEV shedding);
# C2-100 = SinDef pitch + CosDef plunge, LESPcrit 0.18, 100 steps;
# C5-small = ldvm2DOF + delSpalart(30, 10, 1e-6), 120 steps (tests/cases.py c5_2dof).

module UFSRef

import Dierckx: Spline1D, derivative, evaluate
import ForwardDiff
import DelimitedFiles
import Serialization
import NLsolve
import NLsolve: not_in_place
import Statistics: mean
import LinearAlgebra
import Dates

const NL_TRACE = Vector{Any}()

# same call as the reference makes, plus a record of every evaluation point (in call order) and of the result
function traced_nlsolve(f!, x0)
    evals = Vector{Vector{Float64}}()
    function g!(F, x)
        push!(evals, copy(x))
        f!(F, x)
    end
    soln = NLsolve.nlsolve(g!, x0)
    push!(NL_TRACE, Dict("x0" => copy(x0), "evals" => evals, "zero" => copy(soln.zero),
                         "iterations" => soln.iterations, "f_calls" => soln.f_calls, "g_calls" => soln.g_calls,
                         "residual_norm" => soln.residual_norm, "method" => string(soln.method)))
    return soln
end

function load_reference(refdir::String)
    src = joinpath(refdir, "src")
    for rel in ("kinem.jl", "utils.jl", "delVort.jl", "lowOrder2D/typedefs.jl", "lowOrder2D/calcs.jl",
                "lowOrder2D/solvers.jl", "lowOrder2D/postprocess.jl")
        txt = read(joinpath(src, rel), String)
        if rel == "lowOrder2D/solvers.jl"
            occursin("startflag == 0ml", txt) || @warn "solvers.jl: the `0ml` typo was not found (already fixed upstream?)"
            txt = replace(txt, "startflag == 0ml" => "startflag == 0")
            txt = replace(txt, "nlsolve(not_in_place(" => "traced_nlsolve(not_in_place(")
        end
        include_string(UFSRef, txt, joinpath(src, rel))
    end
end

end # module

using Printf

refdir = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia julia_ref.jl <reference checkout> <output dir>")
outdir = length(ARGS) >= 2 ? ARGS[2] : "tests/golden/julia"
mkpath(outdir)
UFSRef.load_reference(refdir)
const U = UFSRef

# ---- minimal JSON writer (no JSON.jl dependency); %.17g round-trips Float64 -----------------------------------
jnum(x::Integer) = string(x)
jnum(x::AbstractFloat) = isfinite(x) ? @sprintf("%.17g", x) : "null"
tojson(x::Number) = jnum(x)
tojson(x::AbstractString) = "\"" * escape_string(x) * "\""
tojson(x::Symbol) = tojson(string(x))
tojson(x::AbstractVector) = "[" * join((tojson(v) for v in x), ",") * "]"
tojson(x::AbstractMatrix) = "[" * join((tojson(collect(x[i, :])) for i in 1:size(x, 1)), ",") * "]"      # row-major list of rows
tojson(x::AbstractDict) = "{" * join((tojson(string(k)) * ":" * tojson(v) for (k, v) in x), ",") * "}"

vorts(v) = Dict("x" => [q.x for q in v], "z" => [q.z for q in v], "s" => [q.s for q in v], "vc" => [q.vc for q in v],
                "vx" => [q.vx for q in v], "vz" => [q.vz for q in v])

function surf_state(surf)
    Dict("bnd_x" => copy(surf.bnd_x), "bnd_z" => copy(surf.bnd_z), "uind" => copy(surf.uind), "wind" => copy(surf.wind),
         "downwash" => copy(surf.downwash), "a0" => surf.a0[1], "a0dot" => surf.a0dot[1], "a0prev" => surf.a0prev[1],
         "aterm" => copy(surf.aterm), "adot" => copy(surf.adot), "aprev" => copy(surf.aprev), "levflag" => Int(surf.levflag[1]),
         "bv" => vorts(surf.bv),
         "kinem" => [surf.kinem.alpha, surf.kinem.h, surf.kinem.alphadot, surf.kinem.hdot, surf.kinem.u, surf.kinem.udot])
end

# run `driver` for nsteps from a fresh state; also the state after 1, 2 and 3 steps (per-step induced velocities)
function run_case(name, mk, driver; nsteps, dtstar, extra = () -> ())
    res = Dict{String,Any}("case" => name, "driver" => string(driver), "nsteps" => nsteps, "dtstar" => dtstar,
                           "julia" => string(VERSION))
    cd(mktempdir()) do
        early = Any[]
        for k in 1:3
            surf, field, args = mk()
            empty!(U.NL_TRACE)
            mat, surf, field = driver(surf, field, args..., k, dtstar, 0, 0, 1000., extra()...)
            push!(early, Dict("steps" => k, "surf" => surf_state(surf), "tev" => vorts(field.tev), "lev" => vorts(field.lev),
                              "mat_row" => collect(mat[end, :])))
        end
        res["early"] = early
        surf, field, args = mk()
        empty!(U.NL_TRACE)
        mat, surf, field = driver(surf, field, args..., nsteps, dtstar, 0, 0, 1000., extra()...)
        res["mat"] = mat
        res["surf"] = surf_state(surf)
        res["tev"] = vorts(field.tev); res["lev"] = vorts(field.lev); res["extv"] = vorts(field.extv)
        res["nlsolve_trace_first_10_calls"] = U.NL_TRACE[1:min(10, length(U.NL_TRACE))]
        res["nlsolve_calls"] = length(U.NL_TRACE)
        if length(args) == 2      # ldvm2DOF: final KinemPar2DOF in field order (typedefs.jl:193-220)
            k2 = args[2]
            res["kinem2dof"] = [getfield(k2, f) for f in fieldnames(typeof(k2))]
        end
    end
    open(joinpath(outdir, name * ".json"), "w") do f
        write(f, tojson(res))
    end
    println("wrote ", joinpath(outdir, name * ".json"), "  Cl_end = ", res["mat"][end, 6])
end

d2r = pi / 180
# C1: test/ldvmLinTest.jl:1-29
mk_c1() = (U.TwoDSurf("FlatPlate", 0.25, U.KinemDef(U.ConstDef(5. * d2r), U.ConstDef(0.), U.ConstDef(1.)), [10.15;], rho = 0.016),
           U.TwoDFlowField(), ())
dt_c1 = U.find_tstep(U.ConstDef(5. * d2r))
n_c1 = Int(round(7. / dt_c1)) + 1
run_case("c1_ldvm", mk_c1, U.ldvm; nsteps = n_c1, dtstar = dt_c1, extra = () -> (U.delNone(),))
run_case("c1_ldvmLin", mk_c1, U.ldvmLin; nsteps = n_c1, dtstar = dt_c1, extra = () -> (U.delNone(),))
run_case("c1_lautat", mk_c1, U.lautat; nsteps = 200, dtstar = dt_c1, extra = () -> (U.delNone(),))
# C1r: LEV-shedding Eldredge ramp (tests/cases.py c1r_surf)
eld = U.EldUpDef(45. * d2r, 0.2, 0.8)
mk_c1r() = (U.TwoDSurf("FlatPlate", 0.0, U.KinemDef(eld, U.ConstDef(0.), U.ConstDef(1.)), [0.11;]), U.TwoDFlowField(), ())
run_case("c1r_ldvm_150", mk_c1r, U.ldvm; nsteps = 150, dtstar = U.find_tstep(eld), extra = () -> (U.delNone(),))
# C2: sinusoidal pitch + cosine plunge (tests/cases.py c2_surf)
sind = U.SinDef(0., 10. * d2r, 0.4, 0.)
mk_c2() = (U.TwoDSurf("FlatPlate", 0.25, U.KinemDef(sind, U.CosDef(0., 0.1, 0.4, 0.), U.ConstDef(1.)), [0.18;]), U.TwoDFlowField(), ())
run_case("c2_ldvm_100", mk_c2, U.ldvm; nsteps = 100, dtstar = U.find_tstep(sind), extra = () -> (U.delNone(),))
# C5-small: ldvm2DOF with Spalart merging (tests/cases.py c5_2dof)
function mk_c5()
    strpar = U.TwoDOFPar(0.2, 0.5, 0.05, 1.0, 0.5, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0)
    kin0 = U.KinemPar2DOF(10. * d2r, 0., 0., 0., 1.)
    surf = U.TwoDSurf("FlatPlate", 0.35, U.KinemDef(U.ConstDef(kin0.alpha), U.ConstDef(0.), U.ConstDef(1.)), [0.2;])
    (surf, U.TwoDFlowField(), (strpar, kin0))
end
run_case("c5_ldvm2dof_120", mk_c5, U.ldvm2DOF; nsteps = 120, dtstar = 0.015, extra = () -> (U.delSpalart(30, 10., 1e-6),))
println("done: commit the files under ", outdir, " and run `pytest tests/test_julia_golden.py`")
