"""CPU tests: the oracle against every pin that exists for this path.

The reference has NO golden vectors; what exists is (a) the bound asserted by
test/ldvmLinTest.jl:33-36, (b) identities that follow from the reference formulas, (c) the
committed fixture tests/golden/c1_ldvm_probe.json, whose values were produced by this oracle
and agree to all printed digits with the surveyor's independent numpy restatement
(SURVEY.md appendix B)."""
import json
import math
import os

import numpy as np
import pytest

from cases import c1_surf, c1r_surf

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_two_vortex_kat(oracle):
    # speed G*d/(2 pi sqrt(vc^4 + d^4)), perpendicular to the separation; 0 at d = 0
    g, vc = 0.7, 0.02
    for d in (0.0, 1e-3, 0.02, 0.5, 30.0):
        u, w = oracle.ind_vel([0.0], [0.0], [g], [vc], [d], [0.0])
        speed = g * d / (2 * math.pi * math.sqrt(vc ** 4 + d ** 4))
        assert abs(u[0]) < 1e-18
        assert abs(w[0] + speed) <= 1e-15 * max(1.0, speed)   # xdist = -d -> w = -speed
    u, w = oracle.ind_vel([0.0], [0.0], [g], [vc], [1e3], [0.0])
    assert abs(-w[0] - g / (2 * math.pi * 1e3)) < 1e-12       # d >> vc -> G/(2 pi d)


def test_mutual_ind_matches_ind_vel_and_antisymmetry(oracle):
    x, z, g, vc = oracle.s_wake(400)
    vx, vz = oracle.mutual_ind(x, z, g, vc)
    u, w = oracle.ind_vel(x, z, g, vc, x, z)       # self term is exactly 0
    scale = np.max(np.abs(u)) + np.max(np.abs(w))
    assert np.max(np.abs(vx - u)) < 1e-13 * scale and np.max(np.abs(vz - w)) < 1e-13 * scale
    # uniform vc: sum_i G_i v_i = 0
    assert abs(np.sum(g * vx)) < 1e-15 and abs(np.sum(g * vz)) < 1e-15
    # accumulation semantics (calcs.jl:502-506)
    vx2, vz2 = oracle.mutual_ind(x, z, g, vc, vx=np.ones(400), vz=-np.ones(400))
    assert np.allclose(vx2 - 1.0, vx, atol=1e-14) and np.allclose(vz2 + 1.0, vz, atol=1e-14)


def test_truth_sum_close_to_reference_order(oracle):
    x, z, g, vc = oracle.s_wake(2000, uniform_vc=False)
    u, w = oracle.ind_vel(x, z, g, vc, x[:50], z[:50])
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, x[:50], z[:50])
    assert np.max(np.abs(u - tu)) / np.max(au) < 1e-13
    assert np.max(np.abs(w - tw)) / np.max(aw) < 1e-13


def test_reference_ldvmlin_test_bound(oracle, uf):
    """test/ldvmLinTest.jl:31-36: |2 pi alpha - Cl_end| < 0.5 * 2 pi alpha after 468 steps"""
    surf, fld, nsteps, dtstar = c1_surf(uf)
    assert nsteps == 468 and dtstar == 0.015
    tab = uf.kinematics_table(surf, fld, nsteps, dtstar)
    mat = oracle.Sim(surf, fld).run(oracle.DRIVER_LDVMLIN, tab, dtstar)
    cl_expected = 2 * math.pi * 5.0 * math.pi / 180
    assert abs(cl_expected - mat[-1, 5]) < 0.5 * cl_expected
    # ldvm (exact mode) agrees with the closed-form linearised solver to ~1e-5 at t* = 7
    mat2 = oracle.Sim(surf, fld).run(oracle.DRIVER_LDVM, tab, dtstar, solver_mode=oracle.SOLVER_EXACT)
    assert abs(mat2[-1, 5] - mat[-1, 5]) < 2e-5
    assert abs(mat2[0, 5] - mat[0, 5]) / mat[0, 5] < 2e-3


def test_golden_c1_probe(oracle, uf):
    gold = json.load(open(os.path.join(GOLD, "c1_ldvm_probe.json")))
    surf, fld, nsteps, dtstar = c1_surf(uf)
    tab = uf.kinematics_table(surf, fld, nsteps, dtstar)
    sim = oracle.Sim(surf, fld)
    mat = sim.run(oracle.DRIVER_LDVM, tab, dtstar, solver_mode=oracle.SOLVER_EXACT)   # the probe values are exact-root ones
    for row in gold["rows"]:
        k = row["step"] - 1
        got = mat[k]
        exp = np.array([row["t"], row["alpha"], row["h"], row["u"], row["a0"], row["cl"], row["cd"], row["cm"]])
        assert np.allclose(got, exp, rtol=1e-10, atol=1e-12), (row["step"], got, exp)
    tev = sim.get_vorts(0)
    assert len(tev["x"]) == gold["ntev"] and len(sim.get_vorts(1)["x"]) == 0
    assert abs(tev["s"].sum() - gold["sum_gamma_tev"]) < 1e-10
    # surveyor's independent numpy restatement, printed digits
    assert abs(mat[-1, 5] - 0.500078179426) < 1e-11 and abs(mat[0, 5] - 9.704962503846) < 1e-11
    assert abs(tev["x"][0] + 0.142135639834) < 1e-11 and abs(tev["s"][0] + 0.041865380402) < 1e-11


def test_kelvin_invariant(oracle, uf):
    """sum G_tev + sum G_lev + uref c pi (A0 + A1/2) = 0 every step (typedefs.jl:249-251)"""
    surf, fld, nsteps, dtstar = c1r_surf(uf)
    tab = uf.kinematics_table(surf, fld, 200, dtstar)
    sim = oracle.Sim(surf, fld)
    for k in range(0, 200, 40):
        sim.run(oracle.DRIVER_LDVM, tab[k:k + 40], dtstar, solver_mode=oracle.SOLVER_EXACT)   # the identities hold at the root
        s = sim.get_surf()
        tot = sim.get_vorts(0)["s"].sum() + sim.get_vorts(1)["s"].sum()
        assert abs(tot + s["uref"] * s["c"] * math.pi * (s["a0"] + s["aterm"][0] / 2)) < 1e-12
    assert len(sim.get_vorts(1)["x"]) > 20      # the ramp does shed LEVs
    # with LEV shedding on, |A0| is clipped to lespcrit at shedding steps
    s = sim.get_surf()
    if s["levflag"] == 1:
        assert abs(abs(s["a0"]) - s["lespcrit"]) < 1e-12


def test_nlsolve_mode_stale_state(oracle, uf):
    """the emulated NLsolve leaves the surf state one finite-difference probe
    away from the root; the strengths themselves agree with the exact solve."""
    surf, fld, nsteps, dtstar = c1_surf(uf)
    tab = uf.kinematics_table(surf, fld, 60, dtstar)
    a, b = oracle.Sim(surf, fld), oracle.Sim(surf, fld)
    m0 = a.run(oracle.DRIVER_LDVM, tab, dtstar, solver_mode=0)
    m1 = b.run(oracle.DRIVER_LDVM, tab, dtstar, solver_mode=1)
    assert np.max(np.abs(a.get_vorts(0)["s"] - b.get_vorts(0)["s"])) < 1e-4
    d = np.max(np.abs(m0[:, 5] - m1[:, 5]))
    assert 1e-5 < d < 5e-2


def test_spalart_merge_and_2dof_run(oracle, uf):
    from cases import c5_2dof
    surf, fld, strpar, kin0 = c5_2dof(uf)
    dt = 0.015
    tab = uf.kinematics_table(surf, fld, 80, dt, two_dof=True)
    sim = oracle.Sim(surf, fld)
    sim.set_2dof(strpar.as_array(), kin0.as_array())
    mat = sim.run(oracle.DRIVER_LDVM2DOF, tab, dt, delvort=(1, 30, 10.0, 1e-6))
    assert np.all(np.isfinite(mat))
    assert len(sim.get_vorts(0)["x"]) < 80          # merging removed some TEVs
    assert abs(mat[-1, 1] - kin0.alpha) > 1e-6      # the structure responded


def test_kinematics_table_matches_oracle_callables(oracle, uf):
    import unsflow_b200 as u
    defs = [(0, [0.3], u.ConstDef(0.3)), (1, [0.1, 0.2, 0.4, 0.3], u.SinDef(0.1, 0.2, 0.4, 0.3)),
            (2, [0.1, 0.2, 0.4, 0.3], u.CosDef(0.1, 0.2, 0.4, 0.3)), (3, [0.7, 0.2, 0.8], u.EldUpDef(0.7, 0.2, 0.8)),
            (4, [0.7, 0.2, 0.8, 2.0], u.EldUptstartDef(0.7, 0.2, 0.8, 2.0)),
            (5, [1.0, 0.0, 2.0, 3.0], u.LinearDef(1.0, 0.0, 2.0, 3.0))]
    for kind, p, d in defs:
        for t in (0.0, 0.5, 1.3, 2.7, 6.0):
            v, dv = oracle.kinem_eval(kind, p, t)
            assert abs(v - d(t)) < 1e-14 and abs(dv - d.deriv(t)) < 1e-13
            if kind != 5:   # analytic derivative vs central difference
                h = 1e-6
                assert abs((d(t + h) - d(t - h)) / (2 * h) - d.deriv(t)) < 1e-7


def test_multi_surface_oracle_consistency(oracle, uf):
    """multi-surface restatement (solvers.jl:270-445): one surface == the single-surface driver;
    two surfaces shed placeholder LEVs (calcs.jl:446) without NaNs; each surface keeps its own
    Kelvin invariant contribution finite"""
    def mk(x0, phase, lesp):
        kin = uf.KinemDef(uf.SinDef(5 * math.pi / 180, 15 * math.pi / 180, 0.6, phase), uf.ConstDef(0.0), uf.ConstDef(1.0))
        return uf.TwoDSurf("FlatPlate", 0.25, kin, [lesp], initpos=(x0, 0.0))
    fld, dt = uf.TwoDFlowField(), 0.012
    tab1 = uf.kinematics_table(mk(0.0, 0.0, 0.12), fld, 30, dt)
    m1 = oracle.MultiSim([mk(0.0, 0.0, 0.12)], fld).run(tab1[:, None, :], dt, solver_mode=oracle.SOLVER_EXACT)
    ms = oracle.Sim(mk(0.0, 0.0, 0.12), fld).run(oracle.DRIVER_LDVM, tab1, dt, solver_mode=oracle.SOLVER_EXACT)
    assert np.max(np.abs(m1 - ms)) < 1e-12
    surfs = [mk(0.0, 0.0, 0.12), mk(1.5, 1.0, 0.2)]
    tabs = np.stack([uf.kinematics_table(s, fld, 120, dt) for s in surfs], axis=1)
    sim = oracle.MultiSim(surfs, fld)
    mat = sim.run(tabs, dt, solver_mode=oracle.SOLVER_EXACT)
    lev = sim.get_vorts(1)
    assert np.all(np.isfinite(mat)) and np.sum(lev["vc"] == 0) > 0 and np.sum(lev["vc"] > 0) > 0
    assert len(sim.get_vorts(0)["x"]) == 240


def test_lve_oracle_consistency(oracle, uf):
    """LVE restatement (LVE.jl:6-302): the steady 5-degree plate approaches the same Wagner value as
    ldvm (Cl(t*=7) ~ 0.50); Kelvin's theorem (last row of the influence matrix, calcs.jl:759)
    keeps  sum(bound) + sum(wake)  constant -- not zero, because the reference never stores the
    strength of the FIRST shed TEV (`if length(IFR_field.tev) > 1`, LVE.jl:160)"""
    def total(sim):
        circ, _ = sim.get_circ()
        return circ[:-1].sum() + sim.get_vorts(0)["s"].sum() + sim.get_vorts(1)["s"].sum()

    kin = uf.KinemDef(uf.ConstDef(5 * math.pi / 180), uf.ConstDef(0.0), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [10.0], rho=1.0)
    ifr = uf.TwoDSurf("FlatPlate", 0.25, kin, [10.0], ndiv=70, camberType="linear")
    tab = uf.lve_table(surf, 468, 0.015)
    sim = oracle.LveSim(surf, ifr)
    mat = sim.run(tab, 468, 0.015)
    assert abs(mat[-1, 5] - 0.5065) < 2e-3 and abs(mat[-1, 1] - 5.0) < 1e-12
    short = oracle.LveSim(surf, ifr)
    short.run(np.vstack([tab[:51], tab[51:52]]), 50, 0.015)
    assert abs(total(sim) - total(short)) < 1e-12 and abs(total(sim)) > 1e-3
    assert sim.get_vorts(0)["s"][0] == 0.0
    kin = uf.KinemDef(uf.EldUpDef(25 * math.pi / 180, 0.2, 0.8), uf.ConstDef(0.0), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [0.2], rho=1.0)
    ifr = uf.TwoDSurf("FlatPlate", 0.25, kin, [0.2], ndiv=70, camberType="linear")
    sim = oracle.LveSim(surf, ifr)
    mat = sim.run(uf.lve_table(surf, 200, 0.015), 200, 0.015)
    circ, gcrit = sim.get_circ()
    assert np.all(np.isfinite(mat)) and len(sim.get_vorts(1)["x"]) > 20
    assert abs(abs(circ[0]) - gcrit) < 2e-4        # leading panel clipped near gamma_crit while shedding


def test_c_oracle_against_independent_numpy_restatement(oracle, uf):
    """oracle/np_restatement.py (vectorised, closed-form affine solves) vs the C oracle (literal
    functors + iterative solve) on LEV-shedding, aeroelastic and vortex-merging runs"""
    from cases import c1r_surf, c5_2dof
    from oracle.np_restatement import NpLdvm
    surf, fld, _, dtstar = c1r_surf(uf)
    n = 110
    tab = uf.kinematics_table(surf, fld, n, dtstar)
    cm = oracle.Sim(surf, fld).run(oracle.DRIVER_LDVM, tab, dtstar, solver_mode=oracle.SOLVER_EXACT)   # the numpy restatement solves in closed form
    npm = NpLdvm(surf, fld)
    nm = npm.run(tab, dtstar)
    assert npm.fam[1].shape[1] > 5                      # LEVs were shed
    assert np.max(np.abs(cm - nm)) < 1e-10
    # ldvm2DOF + Spalart merging
    surf, fld, strpar, kin0 = c5_2dof(uf)
    n = 70
    tab = uf.kinematics_table(surf, fld, n, 0.015, two_dof=True)
    sim = oracle.Sim(surf, fld)
    sim.set_2dof(strpar.as_array(), kin0.as_array())
    cm = sim.run(oracle.DRIVER_LDVM2DOF, tab, 0.015, solver_mode=oracle.SOLVER_EXACT, delvort=(1, 30, 10.0, 1e-6))
    npm = NpLdvm(surf, fld)
    npm.k2 = {k: getattr(kin0, k) for k in kin0.FIELDS}
    npm.sp = dict(zip(("x_alpha", "r_alpha", "kappa", "w_alpha", "w_h", "w_alphadot", "w_hdot", "cubic_h_1", "cubic_h_3",
                       "cubic_alpha_1", "cubic_alpha_3"), strpar.as_array()))
    nm = npm.run(tab, 0.015, delvort=(30, 10.0, 1e-6))
    assert npm.fam[0].shape[1] == len(sim.get_vorts(0)["x"]) < n
    assert np.max(np.abs(cm - nm)) < 1e-10


@pytest.mark.parametrize("driver", ["ldvmLin", "lautat"])
def test_c_oracle_ldvmlin_and_lautat_against_numpy_restatement(oracle, uf, driver):
    """the reference's only tested driver (ldvmLin, test/ldvmLinTest.jl) and lautat, C oracle vs the independent
    numpy restatement: attached flow (config C1) and a LEV-shedding ramp (C1r)"""
    from cases import c1_surf, c1r_surf
    from oracle.np_restatement import NpLdvm
    code = {"ldvmLin": oracle.DRIVER_LDVMLIN, "lautat": oracle.DRIVER_LAUTAT}[driver]
    for mk, n in ((c1_surf, 150), (c1r_surf, 120)):
        r = mk(uf)
        surf, fld, dtstar = r[0], r[1], r[3]
        tab = uf.kinematics_table(surf, fld, n, dtstar, external=(driver != "lautat"))
        cm = oracle.Sim(surf, fld).run(code, tab, dtstar, solver_mode=oracle.SOLVER_EXACT)
        npm = NpLdvm(surf, fld)
        nm = npm.run(tab, dtstar, driver=driver)
        assert np.max(np.abs(cm - nm)) < 1e-10, (driver, mk.__name__, float(np.max(np.abs(cm - nm))))
        if driver == "ldvmLin" and mk is c1r_surf:
            assert npm.fam[1].shape[1] > 5


def test_exact_mode_two_root_step_found_by_the_fuzz_sweep(oracle, uf):
    """tools/fuzz_ldvm.py, seed 7, case 196: at step 55 the Kelvin-Kutta system (LESP target switching on the sign of
    A0, typedefs.jl:340-344) has two self-consistent roots and a Newton iteration started at [-0.01, 0.01] settles
    on A0 = -LESPcrit although the shedding was triggered by A0 > +LESPcrit.  Exact-root mode is defined as the root on
    the branch of the triggering A0: the C oracle (sign pinned during the solve) must agree with the independent
    numpy restatement (closed form on that branch) through and past that step."""
    import importlib.util
    from oracle.np_restatement import NpLdvm
    spec = importlib.util.spec_from_file_location("fuzz_ldvm", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "fuzz_ldvm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(7)
    for _ in range(197):
        cs = mod.make_case(rng)
    assert cs["driver"] == oracle.DRIVER_LDVM and cs["mode"] == 0 and isinstance(cs["delv"], uf.delSpalart)
    om = mod.oracle_history(cs)
    surf, fld, delv = cs["surf"], cs["fld"], cs["delv"]
    dt = cs["dtstar"] * surf.c / surf.uref
    tab = uf.kinematics_table(surf, fld, cs["nsteps"], dt)
    nm = NpLdvm(surf, fld).run(tab, dt, delvort=(delv.limit, delv.dist, delv.tol))
    a0 = om[:, 4]
    assert np.any(np.abs(a0 - cs["lesp"]) < 1e-12) and np.any(np.abs(a0 + cs["lesp"]) < 1e-12)     # sheds on both branches
    assert np.max(np.abs(om - nm)) < 1e-9


def test_delcp_and_edge_velocity_mirror_against_oracle_restatement(oracle, uf):
    """calc_delcp / calc_edgeVel (postprocess.jl:184-244): the host mirror that writes `delcp_edgevel` from a device
    snapshot against the C restatement, on an LEV-shedding state with a non-zero external velocity"""
    from unsflow_b200.postprocess import calc_delcp, calc_edgeVel
    surf, fld, _, dtstar = c1r_surf(uf)
    tab = uf.kinematics_table(surf, fld, 90, dtstar)
    sim = oracle.Sim(surf, fld)
    sim.run(oracle.DRIVER_LDVM, tab, dtstar)
    s = sim.get_surf()
    for name in ("uind", "wind", "aterm", "adot"):
        getattr(surf, name)[:] = s[name]
    surf.a0[0], surf.a0dot[0] = s["a0"], s["a0dot"]
    for a in ("alpha", "h", "alphadot", "hdot", "u", "udot"):
        setattr(surf.kinem, a, s[a])
    vels = [0.07, -0.03]
    p_com, p_in, p_out = calc_delcp(surf, vels)
    qu, ql = calc_edgeVel(surf, vels)
    o = sim.delcp_edgevel(surf.rho, *vels)
    for got, ref, nm in zip((p_com, p_in, p_out, qu, ql), o, ("p_com", "p_in", "p_out", "qu", "ql")):
        assert np.max(np.abs(ref)) > 1e-3, nm
        assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref))), nm
