# UnsflowCUDA.jl -- the reference-side binding of libunsflow_cuda (include/unsflow.h).
#
# Drop-in methods behind the existing UnsteadyFlowSolvers API: `ldvm`, `ldvm2DOF`, `ldvmLin`,
# `lautat` on TwoDSurf / TwoDFlowField / TwoDVort, plus the fine-grained `ind_vel`,
# `mutual_ind`, `wakeroll`, `update_indbound`.  Host code stays Julia; every call below is a thin
# `ccall` into the C ABI.  There is no CPU fallback: a non-zero status becomes `error(msg)`.
#
# NOTE: no `julia` binary exists in the build image, so this file is reviewed, not executed,
# there; tests/ drive exactly the same C symbols through Python ctypes
# (unsteadyflowsolvers.jl_b200/_cabi.py), struct for struct.
#
#   using UnsteadyFlowSolvers; include("UnsflowCUDA.jl"); using .UnsflowCUDA
#   mat, surf, curfield = UnsflowCUDA.ldvm(surf, curfield, nsteps, dtstar)
module UnsflowCUDA

using UnsteadyFlowSolvers
const UFS = UnsteadyFlowSolvers
import UnsteadyFlowSolvers: TwoDVort, TwoDSurf, TwoDFlowField, TwoDOFPar, KinemPar2DOF, delNone, delSpalart
import DelimitedFiles

const LIB = get(ENV, "UNSFLOW_CUDA_LIB", joinpath(@__DIR__, "..", "unsteadyflowsolvers.jl_b200", "libunsflow_cuda.so"))

lasterr() = unsafe_string(ccall((:unsflow_last_error, LIB), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error("libunsflow_cuda: " * lasterr())

# ---- struct mirrors of include/unsflow.h -------------------------------------------------------
struct CVorts            # unsflow_vorts
    n::Int64
    x::Ptr{Cdouble}; z::Ptr{Cdouble}; s::Ptr{Cdouble}; vc::Ptr{Cdouble}; vx::Ptr{Cdouble}; vz::Ptr{Cdouble}
end

struct CSurf             # unsflow_surf
    ndiv::Int32; naterm::Int32
    c::Cdouble; uref::Cdouble; pvt::Cdouble
    lespcrit::Cdouble
    levflag::Int32; _pad::Int32
    kinem::NTuple{6,Cdouble}
    a0::Cdouble; a0dot::Cdouble; a0prev::Cdouble
    theta::Ptr{Cdouble}; x::Ptr{Cdouble}; cam::Ptr{Cdouble}; cam_slope::Ptr{Cdouble}
    bnd_x::Ptr{Cdouble}; bnd_z::Ptr{Cdouble}
    uind::Ptr{Cdouble}; wind::Ptr{Cdouble}; downwash::Ptr{Cdouble}
    aterm::Ptr{Cdouble}; adot::Ptr{Cdouble}; aprev::Ptr{Cdouble}
    bv::CVorts
end

struct CField            # unsflow_field
    u::Cdouble; w::Cdouble
    tev::CVorts; lev::CVorts; extv::CVorts
end

struct COpts             # unsflow_opts
    driver::Int32; solve_mode::Int32; delvort_kind::Int32; device::Int32
    del_limit::Int64; del_dist::Cdouble; del_tol::Cdouble
    dt::Cdouble; max_steps::Int64
end

# Vector{TwoDVort} (array of boxed mutable structs, typedefs.jl:11-18) <-> SoA
struct SoA
    x::Vector{Float64}; z::Vector{Float64}; s::Vector{Float64}; vc::Vector{Float64}; vx::Vector{Float64}; vz::Vector{Float64}
end
SoA(v::Vector{TwoDVort}) = SoA(map(q -> q.x, v), map(q -> q.z, v), map(q -> q.s, v), map(q -> q.vc, v), map(q -> q.vx, v), map(q -> q.vz, v))
SoA(n::Integer) = SoA((zeros(n) for _ = 1:6)...)
cvorts(a::SoA) = CVorts(length(a.x), pointer(a.x), pointer(a.z), pointer(a.s), pointer(a.vc), pointer(a.vx), pointer(a.vz))
function unpack!(v::Vector{TwoDVort}, a::SoA, n::Integer)
    resize!(v, 0)
    for i = 1:n
        push!(v, TwoDVort(a.x[i], a.z[i], a.s[i], a.vc[i], a.vx[i], a.vz[i]))
    end
    v
end

# ---- fine-grained drop-ins (calcs.jl) ------------------------------------------------------------
"ind_vel(src, t_x, t_z) -> (uind, wind)   replaces calcs.jl:462-477"
function ind_vel(src::Vector{TwoDVort}, t_x::Vector{Float64}, t_z::Vector{Float64})
    a = SoA(src); uind = zeros(length(t_x)); wind = zeros(length(t_x))
    check(ccall((:unsflow_ind_vel, LIB), Cint,
        (Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        length(src), a.x, a.z, a.s, a.vc, length(t_x), t_x, t_z, uind, wind))
    uind, wind
end
function ind_vel(src::TwoDVort, t_x, t_z)            # calcs.jl:480-488: broadcasts, so matrix-shaped targets (a meshgrid) keep their shape
    u, w = ind_vel([src], vec(collect(Float64, t_x)), vec(collect(Float64, t_z)))
    reshape(u, size(t_x)), reshape(w, size(t_x))
end
"velocity induced by ALL of `vorts` at mesh points in one device call (the per-vortex loop of LVE.jl:176-183)"
function ind_vel(src::Vector{TwoDVort}, t_x::Matrix{Float64}, t_z::Matrix{Float64})
    u, w = ind_vel(src, vec(t_x), vec(t_z))
    reshape(u, size(t_x)), reshape(w, size(t_x))
end

"mutual_ind(vorts) -> vorts (accumulates into vx, vz)   replaces calcs.jl:491-510"
function mutual_ind(vorts::Vector{TwoDVort})
    a = SoA(vorts)
    check(ccall((:unsflow_mutual_ind, LIB), Cint,
        (Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        length(vorts), a.x, a.z, a.s, a.vc, a.vx, a.vz))
    for i = 1:length(vorts)
        vorts[i].vx = a.vx[i]; vorts[i].vz = a.vz[i]
    end
    vorts
end

"update_indbound(surf, curfield) -> surf   replaces calcs.jl:35-38"
function update_indbound(surf::TwoDSurf, curfield::TwoDFlowField)
    a = SoA([curfield.tev; curfield.lev; curfield.extv])
    uind = zeros(surf.ndiv); wind = zeros(surf.ndiv)
    check(ccall((:unsflow_update_indbound, LIB), Cint,
        (Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        length(a.x), a.x, a.z, a.s, a.vc, surf.ndiv, surf.bnd_x, surf.bnd_z, uind, wind))
    surf.uind[1:surf.ndiv] = uind; surf.wind[1:surf.ndiv] = wind
    surf
end

"wakeroll(surf, curfield, dt) -> curfield   replaces calcs.jl:222-294"
function wakeroll(surf::TwoDSurf, curfield::TwoDFlowField, dt)
    all = [curfield.tev; curfield.lev; curfield.extv]      # same order as calcs.jl:245 (shares the boxes)
    a = SoA(all); b = SoA(surf.bv)
    check(ccall((:unsflow_wakeroll, LIB), Cint,
        (Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
         Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Cdouble),
        length(all), a.x, a.z, a.s, a.vc, a.vx, a.vz, length(surf.bv), b.x, b.z, b.s, b.vc,
        curfield.u[1], curfield.w[1], dt))
    for i = 1:length(all)
        all[i].x = a.x[i]; all[i].z = a.z[i]; all[i].vx = a.vx[i]; all[i].vz = a.vz[i]
    end
    curfield
end

# ---- device-resident drivers (solvers.jl) --------------------------------------------------------
# host-tabulated prescribed motion: rows [t, alpha, h, alphadot, hdot, u, udot, U, W] at the
# running-sum times t = t + dt (solvers.jl:180) using the reference's own update_kinem /
# update_externalvel (calcs.jl:75-198) on scratch copies
function kinematics_table(surf::TwoDSurf, curfield::TwoDFlowField, nsteps, dt, t0; external = true, twodof = false)
    tab = zeros(9, nsteps)                      # column per step == row-major nsteps x 9 for C
    s = deepcopy(surf); f = deepcopy(curfield); t = t0
    for i = 1:nsteps
        t = t + dt
        external && UFS.update_externalvel(f, t)
        twodof || UFS.update_kinem(s, t)
        k = s.kinem
        tab[:, i] = [t, k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot, f.u[1], f.w[1]]
    end
    tab
end

function run_driver(driver::Integer, surf::TwoDSurf, curfield::TwoDFlowField, nsteps::Int64, dtstar::Float64, startflag,
                    writeflag, writeInterval, delvort; maxwrite = 50, nround = 6, solve_mode = 1, strpar = nothing, kinem = nothing)
    startflag == 0 || throw("invalid start flag, should be 0 or 1")        # solvers.jl:157 (restart is broken upstream)
    t = 0.0
    dt = dtstar * surf.c / surf.uref                                       # solvers.jl:161
    tab = kinematics_table(surf, curfield, nsteps, dt, t; external = driver != 3, twodof = driver == 1)
    tev = SoA(curfield.tev); lev = SoA(curfield.lev); extv = SoA(curfield.extv); bv = SoA(surf.bv)
    k = surf.kinem
    mk_surf() = CSurf(surf.ndiv, surf.naterm, surf.c, surf.uref, surf.pvt, surf.lespcrit[1], surf.levflag[1], 0,
        (k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot), surf.a0[1], surf.a0dot[1], surf.a0prev[1],
        pointer(surf.theta), pointer(surf.x), pointer(surf.cam), pointer(surf.cam_slope), pointer(surf.bnd_x), pointer(surf.bnd_z),
        pointer(surf.uind), pointer(surf.wind), pointer(surf.downwash), pointer(surf.aterm), pointer(surf.adot), pointer(surf.aprev),
        cvorts(bv))
    kind, lim, dist, tol = typeof(delvort) == delSpalart ? (1, delvort.limit, delvort.dist, delvort.tol) : (0, 0, 0.0, 0.0)
    opts = Ref(COpts(driver, solve_mode, kind, -1, lim, dist, tol, dt, nsteps))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mat = Matrix{Float64}(undef, nsteps, 8)
    GC.@preserve surf tev lev extv bv tab mat begin
        cs = Ref(mk_surf()); cf = Ref(CField(curfield.u[1], curfield.w[1], cvorts(tev), cvorts(lev), cvorts(extv)))
        check(ccall((:unsflow_ldvm_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{CSurf}, Ptr{CField}, Ptr{COpts}), h, cs, cf, opts))
        try
            if driver == 1
                sp = [getfield(strpar, f) for f in fieldnames(TwoDOFPar)]
                kk = [getfield(kinem, f) for f in fieldnames(KinemPar2DOF)]
                check(ccall((:unsflow_ldvm_set_2dof, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h[], sp, kk))
            end
            check(ccall((:unsflow_ldvm_set_kinematics, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), h[], nsteps, tab))
            # write schedule of solvers.jl:164-175: run up to each stamp, download, writeStamp (postprocess.jl:34-113)
            stamps = writeflag == 1 ? [Int(round(writeInterval * i / dt)) for i = 1:maxwrite] : Int[]
            done = 0
            for stop in sort(unique([filter(s -> 0 < s <= nsteps, stamps); nsteps]))
                if stop > done
                    sub = Matrix{Float64}(undef, stop - done, 8)
                    check(ccall((:unsflow_ldvm_run, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), h[], stop - done, sub))
                    mat[done+1:stop, :] = sub
                    done = stop
                end
                if stop in stamps
                    download!(h[], surf, curfield)
                    UFS.writeStamp("$(round(tab[1, stop], sigdigits=nround))", tab[1, stop], surf, curfield)
                end
            end
            download!(h[], surf, curfield)
            if driver == 1
                kk = zeros(22)
                check(ccall((:unsflow_ldvm_get_2dof, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h[], kk))
                for (i, f) in enumerate(fieldnames(KinemPar2DOF))
                    setfield!(kinem, f, kk[i])
                end
            end
        finally
            ccall((:unsflow_ldvm_destroy, LIB), Cint, (Ptr{Cvoid},), h[])
        end
    end
    f = open("resultsSummary", "w")                                         # solvers.jl:261-264
    println(f, "#time \t", "alpha(deg) \t", "h/c \t", "u/uref \t", "A0 \t", "Cl \t", "Cd \t", "Cm ")
    DelimitedFiles.writedlm(f, mat)
    close(f)
    mat, surf, curfield
end

function download!(h::Ptr{Cvoid}, surf::TwoDSurf, curfield::TwoDFlowField)
    nt = Ref{Int64}(0); nl = Ref{Int64}(0); ne = Ref{Int64}(0)
    check(ccall((:unsflow_ldvm_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h, nt, nl, ne))
    tev = SoA(nt[]); lev = SoA(nl[]); extv = SoA(ne[]); bv = SoA(surf.ndiv - 1)
    k = surf.kinem
    GC.@preserve surf tev lev extv bv begin
        cs = Ref(CSurf(surf.ndiv, surf.naterm, surf.c, surf.uref, surf.pvt, surf.lespcrit[1], surf.levflag[1], 0,
            (k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot), 0.0, 0.0, 0.0,
            pointer(surf.theta), pointer(surf.x), pointer(surf.cam), pointer(surf.cam_slope), pointer(surf.bnd_x), pointer(surf.bnd_z),
            pointer(surf.uind), pointer(surf.wind), pointer(surf.downwash), pointer(surf.aterm), pointer(surf.adot), pointer(surf.aprev),
            cvorts(bv)))
        cf = Ref(CField(0.0, 0.0, cvorts(tev), cvorts(lev), cvorts(extv)))
        check(ccall((:unsflow_ldvm_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{CSurf}, Ptr{CField}), h, cs, cf))
        s = cs[]; f = cf[]
        surf.levflag[1] = s.levflag
        k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot = s.kinem
        surf.a0[1] = s.a0; surf.a0dot[1] = s.a0dot; surf.a0prev[1] = s.a0prev
        curfield.u[1] = f.u; curfield.w[1] = f.w
    end
    unpack!(curfield.tev, tev, nt[]); unpack!(curfield.lev, lev, nl[]); unpack!(curfield.extv, extv, ne[])
    for i = 1:surf.ndiv-1
        surf.bv[i].s = bv.s[i]; surf.bv[i].x = bv.x[i]; surf.bv[i].z = bv.z[i]
    end
end

"ldvm(surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort; maxwrite, nround)  replaces solvers.jl:144-268"
ldvm(surf::TwoDSurf, curfield::TwoDFlowField, nsteps::Int64 = 500, dtstar::Float64 = 0.015, startflag = 0, writeflag = 0,
     writeInterval = 1000., delvort = delNone(); maxwrite = 50, nround = 6, solve_mode = 1) =
    run_driver(0, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort; maxwrite = maxwrite, nround = nround, solve_mode = solve_mode)

"ldvm2DOF(surf, curfield, strpar, kinem, ...)  replaces solvers.jl:657-788"
ldvm2DOF(surf::TwoDSurf, curfield::TwoDFlowField, strpar::TwoDOFPar, kinem::KinemPar2DOF, nsteps::Int64 = 500, dtstar::Float64 = 0.015,
         startflag = 0, writeflag = 0, writeInterval = 1000., delvort = delNone(); maxwrite = 50, nround = 6, solve_mode = 1) =
    run_driver(1, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort; maxwrite = maxwrite, nround = nround,
               solve_mode = solve_mode, strpar = strpar, kinem = kinem)

"ldvmLin(...)  replaces solvers.jl:448-655"
ldvmLin(surf::TwoDSurf, curfield::TwoDFlowField, nsteps::Int64 = 500, dtstar::Float64 = 0.015, startflag = 0, writeflag = 0,
        writeInterval = 1000., delvort = delNone(); maxwrite = 50, nround = 6) =
    run_driver(2, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort; maxwrite = maxwrite, nround = nround)

"lautat(...)  replaces solvers.jl:42-142 (wakerollup = 1)"
lautat(surf::TwoDSurf, curfield::TwoDFlowField, nsteps::Int64 = 500, dtstar::Float64 = 0.015, startflag = 0, writeflag = 0,
       writeInterval = 1000., delvort = delNone(); maxwrite = 50, nround = 6, solve_mode = 1) =
    run_driver(3, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort; maxwrite = maxwrite, nround = nround, solve_mode = solve_mode)

# ---- multi-surface ldvm(surf::Vector{TwoDSurf}, ...)  replaces solvers.jl:270-445 -------------------
function csurf(surf::TwoDSurf, bv::SoA)
    k = surf.kinem
    CSurf(surf.ndiv, surf.naterm, surf.c, surf.uref, surf.pvt, surf.lespcrit[1], surf.levflag[1], 0,
        (k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot), surf.a0[1], surf.a0dot[1], surf.a0prev[1],
        pointer(surf.theta), pointer(surf.x), pointer(surf.cam), pointer(surf.cam_slope), pointer(surf.bnd_x), pointer(surf.bnd_z),
        pointer(surf.uind), pointer(surf.wind), pointer(surf.downwash), pointer(surf.aterm), pointer(surf.adot), pointer(surf.aprev),
        cvorts(bv))
end

function ldvm(surf::Vector{TwoDSurf}, curfield::TwoDFlowField, nsteps::Int64 = 500, dtstar::Float64 = 0.015, startflag = 0,
              writeflag = 0, writeInterval = 1000., delvort = delNone(); maxwrite = 50, nround = 6, solve_mode = 1)
    startflag == 0 || throw("invalid start flag, should be 0 or 1")
    nsurf = length(surf)
    dt = dtstar * minimum(map(q -> q.c / q.uref, surf))                     # solvers.jl:289
    # table: 9 x nsurf x nsteps column-major == nsteps x nsurf x 9 row-major for C
    tab = zeros(9, nsurf, nsteps)
    for (is, sf) in enumerate(surf)
        tab[:, is, :] = kinematics_table(sf, curfield, nsteps, dt, 0.0)
    end
    tev = SoA(curfield.tev); lev = SoA(curfield.lev); extv = SoA(curfield.extv); bvs = [SoA(sf.bv) for sf in surf]
    kind, lim, dist, tol = typeof(delvort) == delSpalart ? (1, delvort.limit, delvort.dist, delvort.tol) : (0, 0, 0.0, 0.0)
    opts = Ref(COpts(0, solve_mode, kind, -1, lim, dist, tol, dt, nsteps))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mat = Matrix{Float64}(undef, nsteps, 1 + 7 * nsurf)
    GC.@preserve surf tev lev extv bvs tab mat begin
        cs = [csurf(surf[i], bvs[i]) for i = 1:nsurf]
        cf = Ref(CField(curfield.u[1], curfield.w[1], cvorts(tev), cvorts(lev), cvorts(extv)))
        check(ccall((:unsflow_ldvm_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Int32, Ptr{CSurf}, Ptr{CField}, Ptr{COpts}), h, nsurf, cs, cf, opts))
        try
            check(ccall((:unsflow_ldvm_set_kinematics, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), h[], nsteps, tab))
            check(ccall((:unsflow_ldvm_run, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), h[], nsteps, mat))
            nt = Ref{Int64}(0); nl = Ref{Int64}(0); ne = Ref{Int64}(0)
            check(ccall((:unsflow_ldvm_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h[], nt, nl, ne))
            tev = SoA(nt[]); lev = SoA(nl[]); extv = SoA(ne[])
            cs = [csurf(surf[i], bvs[i]) for i = 1:nsurf]
            cf = Ref(CField(0.0, 0.0, cvorts(tev), cvorts(lev), cvorts(extv)))
            GC.@preserve tev lev extv begin
                check(ccall((:unsflow_ldvm_multi_get_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{CSurf}, Ptr{CField}), h[], nsurf, cs, cf))
            end
            for i = 1:nsurf
                sf = surf[i]; c = cs[i]; k = sf.kinem
                sf.levflag[1] = c.levflag
                k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot = c.kinem
                sf.a0[1] = c.a0; sf.a0dot[1] = c.a0dot; sf.a0prev[1] = c.a0prev
                for j = 1:sf.ndiv-1
                    sf.bv[j].s = bvs[i].s[j]; sf.bv[j].x = bvs[i].x[j]; sf.bv[j].z = bvs[i].z[j]
                end
            end
            curfield.u[1] = cf[].u; curfield.w[1] = cf[].w
            unpack!(curfield.tev, tev, nt[]); unpack!(curfield.lev, lev, nl[]); unpack!(curfield.extv, extv, ne[])
        finally
            ccall((:unsflow_ldvm_destroy, LIB), Cint, (Ptr{Cvoid},), h[])
        end
    end
    f = open("resultsSummary", "w")                                         # solvers.jl:438-441
    println(f, "#time \t", "alpha(deg) \t", "h/c \t", "u/uref \t", "A0 \t", "Cl \t", "Cd \t", "Cm ")
    DelimitedFiles.writedlm(f, mat)
    close(f)
    mat, surf, curfield
end

# ---- LVE(surf, curfield, nsteps, dtstar, ...)  replaces the time loop of LVE.jl:6-302 ---------------
# (plotting / animation outputs are not produced: frames = 0, anim and wflag must be false)
function LVE(surf::TwoDSurf, curfield::TwoDFlowField, nsteps::Int64 = 500, dtstar::Float64 = 0.015, frameCnt = 20, view = "square",
             LEV_Stop = 1000, anim = false, animStep = 30, wflag = false, writeInterval = 50)
    (anim || wflag) && error("UnsflowCUDA.LVE: anim / wflag are host-side plotting, use UnsteadyFlowSolvers.LVE")
    IFR_surf = TwoDSurf(surf.coord_file, surf.pvt, surf.kindef, surf.lespcrit; ndiv = surf.ndiv, camberType = "linear")  # LVE.jl:17
    IFR_field = TwoDFlowField()
    ds = sqrt.((surf.x[1:end-1] - surf.x[2:end]).^2 + (surf.cam[1:end-1] - surf.cam[2:end]).^2)                          # :29-41
    cam_slope = asind.((surf.cam[2:end] - surf.cam[1:end-1]) ./ ds)
    vor_loc = hcat(surf.x[1:end-1] + .25 .* ds .* cosd.(cam_slope), surf.cam[1:end-1] + .25 .* ds .* sind.(cam_slope))
    UFS.refresh_vortex(surf, vor_loc)                                                                                    # :55
    # (nsteps + 2) rows [t, alpha, h, alphadot, hdot, u, udot, X0, Z0]: iterations i = 0..nsteps and t + dtstar of the last
    tab = zeros(9, nsteps + 2); s = deepcopy(surf); t = -dtstar
    for i = 1:nsteps+2
        t += dtstar
        UFS.update_kinem(s, t); k = s.kinem
        tab[:, i] = [t, s.kindef.alpha(t), k.h, k.alphadot, k.hdot, k.u, k.udot, -s.kindef.u(t) * t, s.kindef.h(t)]
    end
    bv = SoA(surf.bv); e = SoA(0)
    opts = Ref(COpts(0, 0, 0, -1, 0, 0.0, 0.0, dtstar, nsteps + 1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    m9 = Matrix{Float64}(undef, nsteps + 1, 9)
    circ = zeros(surf.ndiv); gcrit = Ref{Cdouble}(0.0)
    tev = SoA(0); lev = SoA(0)
    GC.@preserve surf bv e tab m9 circ begin
        cs = Ref(csurf(surf, bv)); cf = Ref(CField(0.0, 0.0, cvorts(e), cvorts(e), cvorts(e)))
        check(ccall((:unsflow_lve_create, LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Ptr{CSurf}, Ptr{CField}, Ptr{COpts}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble),
            h, cs, cf, opts, surf.rho, LEV_Stop, IFR_surf.x[end], IFR_surf.cam[end], IFR_surf.kinem.u, IFR_surf.kinem.hdot))
        try
            check(ccall((:unsflow_ldvm_set_kinematics, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), h[], nsteps + 2, tab))
            check(ccall((:unsflow_ldvm_run, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), h[], nsteps + 1, m9))
            nt = Ref{Int64}(0); nl = Ref{Int64}(0); ne = Ref{Int64}(0)
            check(ccall((:unsflow_ldvm_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h[], nt, nl, ne))
            tev = SoA(nt[]); lev = SoA(nl[]); ibv = SoA(surf.ndiv - 1)
            bx = copy(surf.bnd_x); bz = copy(surf.bnd_z)
            GC.@preserve tev lev ibv begin
                cs2 = Ref(csurf(surf, ibv)); cf2 = Ref(CField(0.0, 0.0, cvorts(tev), cvorts(lev), cvorts(e)))
                check(ccall((:unsflow_ldvm_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{CSurf}, Ptr{CField}), h[], cs2, cf2))
                k = surf.kinem; c2 = cs2[]
                k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot = c2.kinem
                surf.levflag[1] = c2.levflag; surf.a0[1] = c2.a0
            end
            surf.bnd_x[:] = bx; surf.bnd_z[:] = bz                       # LVE never touches surf.bnd_x/z
            check(ccall((:unsflow_lve_get_circ, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h[], circ, gcrit))
            unpack!(IFR_field.tev, tev, nt[]); unpack!(IFR_field.lev, lev, nl[])
        finally
            ccall((:unsflow_ldvm_destroy, LIB), Cint, (Ptr{Cvoid},), h[])
        end
    end
    for j = 1:surf.ndiv-1
        surf.bv[j].s = circ[j]                                              # LVE.jl:156
    end
    # curfield = body-frame twin of IFR_field at t + dtstar (LVE.jl:238-242)
    resize!(curfield.tev, 0); resize!(curfield.lev, 0)
    for q in IFR_field.tev; push!(curfield.tev, TwoDVort(0, 0, q.s, 0.02 * surf.c, 0., 0.)); end
    for q in IFR_field.lev; push!(curfield.lev, TwoDVort(0, 0, q.s, 0.02 * surf.c, 0., 0.)); end
    UFS.vor_BFR(surf, tab[1, end], IFR_field, curfield)
    mat = m9[2:end, 1:8]
    TEVhist = hcat(tab[1, 1:nsteps+1], map(q -> q.s, IFR_field.tev)[1:nsteps+1])
    shed = findall(m9[:, 9] .> 0.5)
    LEVhist = hcat(tab[1, shed], map(q -> q.s, IFR_field.lev)[1:length(shed)])
    return 0, IFR_field, mat, TEVhist, LEVhist
end

# ---- ensembles of independent runs (no reference equivalent: a `for` loop over ldvm / ldvm2DOF calls) ----
# `surfs[m]` is member m's TwoDSurf as the user would build it for that loop (same geometry and
# discretisation, own kindef and lespcrit); every member starts from surfs[1]'s state with an empty wake.
# One thread block per member runs the whole time loop on the GPU.  Returns the members' `mat`
# matrices (nsteps x 8 each) and their TEV / LEV counts.
function member_tables(surfs::Vector{TwoDSurf}, nsteps, dt; twodof = false)
    tabs = zeros(9, nsteps, length(surfs))
    for (m, sf) in enumerate(surfs)
        tabs[:, :, m] = kinematics_table(sf, TwoDFlowField(), nsteps, dt, 0.0; twodof = twodof)
    end
    tabs
end

function ldvm_batch(surfs::Vector{TwoDSurf}, nsteps::Int64 = 500, dtstar::Float64 = 0.015, delvort = delNone(); solve_mode = 1)
    surf = surfs[1]; M = length(surfs); dt = dtstar * surf.c / surf.uref
    tabs = member_tables(surfs, nsteps, dt)
    lesp = [sf.lespcrit[1] for sf in surfs]
    bv = SoA(surf.bv); mats = zeros(nsteps, 8, M); nt = zeros(Int64, M); nl = zeros(Int64, M)
    kind, lim, dist, tol = typeof(delvort) == delSpalart ? (1, delvort.limit, delvort.dist, delvort.tol) : (0, 0, 0.0, 0.0)
    GC.@preserve surf bv begin
        cs = Ref(csurf(surf, bv)); opts = Ref(COpts(0, solve_mode, kind, -1, lim, dist, tol, dt, nsteps))
        check(ccall((:unsflow_ldvm_batch_run, LIB), Cint,
            (Ptr{CSurf}, Ptr{COpts}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}),
            cs, opts, M, nsteps, tabs, lesp, mats, nt, nl))
    end
    [mats[:, :, m] for m = 1:M], nt, nl
end

# ldvm2DOF members: strpars[m]::TwoDOFPar, kinems[m]::KinemPar2DOF (mutated to the final state)
function ldvm2DOF_batch(surfs::Vector{TwoDSurf}, strpars::Vector{TwoDOFPar}, kinems::Vector{KinemPar2DOF},
                        nsteps::Int64 = 500, dtstar::Float64 = 0.015, delvort = delNone(); solve_mode = 1)
    surf = surfs[1]; M = length(surfs); dt = dtstar * surf.c / surf.uref
    tabs = member_tables(surfs, nsteps, dt; twodof = true)
    lesp = [sf.lespcrit[1] for sf in surfs]
    sp = [Float64(getfield(strpars[m], f)) for f in fieldnames(TwoDOFPar), m = 1:M]          # 11 x M
    kk = [Float64(getfield(kinems[m], f)) for f in fieldnames(KinemPar2DOF), m = 1:M]       # 22 x M
    kout = similar(kk)
    bv = SoA(surf.bv); mats = zeros(nsteps, 8, M); nt = zeros(Int64, M); nl = zeros(Int64, M)
    kind, lim, dist, tol = typeof(delvort) == delSpalart ? (1, delvort.limit, delvort.dist, delvort.tol) : (0, 0, 0.0, 0.0)
    GC.@preserve surf bv begin
        cs = Ref(csurf(surf, bv)); opts = Ref(COpts(1, solve_mode, kind, -1, lim, dist, tol, dt, nsteps))
        check(ccall((:unsflow_ldvm2dof_batch_run, LIB), Cint,
            (Ptr{CSurf}, Ptr{COpts}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}),
            cs, opts, M, nsteps, tabs, lesp, sp, kk, mats, kout, nt, nl))
    end
    for m = 1:M, (i, f) in enumerate(fieldnames(KinemPar2DOF))
        setfield!(kinems[m], f, kout[i, m])
    end
    [mats[:, :, m] for m = 1:M], nt, nl
end

end # module
