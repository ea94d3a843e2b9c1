"""CPU tests of the host-side mirror of the reference API (types, kinematics tables)."""
import math

import numpy as np

from cases import c1_surf


def test_twodsurf_constructor_matches_reference_formulas(uf):
    surf, fld, nsteps, dtstar = c1_surf(uf)
    assert surf.ndiv == 70 and surf.naterm == 35 and len(surf.bv) == 69          # typedefs.jl:68
    assert abs(surf.theta[-1] - math.pi) < 1e-15 and abs(surf.x[-1] - 1.0) < 1e-15  # :77-82
    a = 5 * math.pi / 180
    # :156-159 -- TE at the origin side, LE at -(c - pvt c) - pvt c cos(alpha)
    assert abs(surf.bnd_x[0] + (0.75 + 0.25 * math.cos(a))) < 1e-15
    assert abs(surf.bnd_z[0] - 0.25 * math.sin(a)) < 1e-15
    assert np.all(surf.bv.vc == 0.02) and surf.levflag[0] == 0
    assert len(fld.tev) == 0 and fld.u[0] == 0


def test_vortlist_semantics(uf):
    v = uf.VortList()
    for i in range(100):
        v.append(uf.TwoDVort(i, -i, 0.1 * i, 0.02))
    assert len(v) == 100 and v[99].x == 99 and v[-1].z == -99
    v[3].s = 7.0
    assert v.s[3] == 7.0
    first = v.popfirst()
    assert first.x == 0 and len(v) == 99 and v[0].x == 1 and v[2].s == 7.0


def test_kinematics_table_running_sum(uf):
    surf, fld, nsteps, dtstar = c1_surf(uf)
    tab = uf.kinematics_table(surf, fld, 10, dtstar)
    t = 0.0
    for i in range(10):
        t = t + dtstar                     # solvers.jl:180, not (i+1)*dt
        assert tab[i, 0] == t
    assert np.all(tab[:, 1] == 5 * math.pi / 180) and np.all(tab[:, 5] == 1.0) and np.all(tab[:, 3] == 0.0)


def test_find_tstep(uf):
    assert uf.find_tstep(uf.ConstDef(1.0)) == 0.015
    assert abs(uf.find_tstep(uf.EldUpDef(0.5, 0.4, 0.8)) - 0.0075) < 1e-15
    assert abs(uf.find_tstep(uf.SinDef(0, 0.5, 0.8, 0)) - 0.015 * 0.2 / 0.4) < 1e-15


def test_writestamp_layout_cpu(uf, tmp_path, monkeypatch):
    """postprocess.jl:34-113 file layout from a host-side state (no GPU involved)"""
    import os
    monkeypatch.chdir(tmp_path)
    surf, fld, _, _ = c1_surf(uf)
    surf.a0[0], surf.a0dot[0] = 0.05, 0.01
    surf.aterm[:] = 0.01 / (1 + np.arange(35))
    fld.tev.append(uf.TwoDVort(0.1, 0.0, -0.02, 0.02))
    uf.writeStamp("0.015", 0.015, surf, fld)
    d = os.path.join("Step Files", "0.015")
    assert sorted(os.listdir(d)) == ["FourierCoeffs", "boundv", "delcp_edgevel", "lev", "tev", "timeKinem"]
    assert np.loadtxt(os.path.join(d, "boundv")).shape == (69, 3)
    assert np.loadtxt(os.path.join(d, "tev")).reshape(-1, 3).shape == (1, 3)
    dc = np.loadtxt(os.path.join(d, "delcp_edgevel"))
    assert dc.shape == (70, 6) and np.all(np.isfinite(dc))
    qu, ql = uf.calc_edgeVel(surf, [0.0, 0.0])
    assert abs(qu[0] - 0.05 / math.sqrt(0.5 * 0.016)) < 1e-14          # postprocess.jl:230


def test_meshgrid_geometry(uf):
    """meshgrid, typedefs.jl:422-512: window, mesh orientation, wake filter (host-only part)"""
    surf = uf.TwoDSurf("FlatPlate", 0.25, uf.KinemDef(uf.ConstDef(0.1), uf.ConstDef(0.2), uf.ConstDef(1.0)), [10.15])
    tev = uf.VortList.from_arrays([0.4, 3.0, -1.5], [0.1, 0.0, 0.2], [1, 1, 1], [0.02] * 3)
    m = uf.meshgrid(surf, tev, uf.VortList(), 0.25, 0.5, 100, "square")
    assert m.x.shape == m.z.shape == (100, 100)
    step = 1.5 / 99
    assert abs(m.x[0, 1] - m.x[0, 0] - step) < 1e-14 and abs(m.z[0, 0] - m.z[1, 0] - step) < 1e-14
    assert abs(m.x[0, 0] - (-0.25 - 0.5 - 0.25)) < 1e-14           # nearBnd + X0 - pvt*c
    assert abs(m.z[0, 0] - (0.75 + 0.2)) < 1e-14                    # top row = uppZ
    assert m.tev.shape == (1, 2) and m.tev[0, 0] == 0.4 and m.lev.shape == (0, 2)
    assert np.all(m.uMat == 1.0) and np.all(m.wMat == 0.0) and np.all(m.velMag == 0.0)
    assert abs(m.camX[0] - ((0 - 0.25) * np.cos(0.1) - 0.5 + 0.25)) < 1e-14
    w = uf.meshgrid(surf, tev, uf.VortList(), 0.25, 0.5, 100, "wake")
    assert w.x.shape == w.z.shape == (55, 100) and w.tev.shape[0] == 1


def test_eldupint_motion_definitions(uf):
    """EldUpIntDef / EldUpInttstartDef (kinem.jl:179-249): value = running rectangle-rule integral of the Eldredge ramp
    with the reference's re-derived step (dt = t / (round(t/0.015 + 1) - 1)); derivative = what ForwardDiff returns
    for that callable (checked by a finite difference inside a window where the step count does not change);
    used for the plunge in update_kinem (calcs.jl:137-142) and, EldUpIntDef only, in the constructor (typedefs.jl:121)"""
    import math
    for m in (uf.EldUpIntDef(0.5, 0.2, 0.8), uf.EldUpInttstartDef(0.3, 0.4, 0.6, 0.5)):
        assert m(0.0) == 0.0 and m.deriv(0.0) == 0.0 and m(0.004) == 0.0            # nsteps == 1 -> literal 0
        for t in (0.3, 1.234, 2.71):
            n = int(round(t / 0.015 + 1))
            dt = t / (n - 1)
            sm = math.pi ** 2 * m.K / (2 * m.amp * (1 - m.a))
            t1 = m._tstart()
            t2 = t1 + m.amp / (2 * m.K)
            ref = sum(((m.K / sm) * math.log(math.cosh(sm * (i * dt - t1)) / math.cosh(sm * (i * dt - t2))) + m.amp / 2) * dt for i in range(n))
            assert abs(m(t) - ref) < 1e-14
            h = 1e-6
            assert int(round((t + h) / 0.015 + 1)) == n == int(round((t - h) / 0.015 + 1))
            assert abs(m.deriv(t) - (m(t + h) - m(t - h)) / (2 * h)) < 1e-8
    # wired into the plunge branch of update_kinem and of the TwoDSurf constructor
    kin = uf.KinemDef(uf.ConstDef(0.1), uf.EldUpIntDef(0.5, 0.2, 0.8), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [0.2], c=2.0, uref=1.5)
    tab = uf.kinematics_table(surf, uf.TwoDFlowField(), 100, 0.015 * 2.0 / 1.5)
    t = tab[-1, 0]
    assert abs(tab[-1, 2] - kin.h(t) * 2.0) < 1e-15 and abs(tab[-1, 4] - kin.h.deriv(t) * 1.5) < 1e-15 and tab[-1, 2] > 0


def test_ensemble_shard_is_a_balanced_partition():
    """ensemble_shard deals the members over the ranks: every member exactly once, shard sizes within one, and the
    expensive members (low LESPcrit, fast pitching) spread evenly instead of all landing on rank 0"""
    import unsflow_b200 as uf
    rng = np.random.default_rng(3)
    nm, ns = 101, 40
    tables = np.zeros((nm, ns, 9))
    tables[:, :, 5] = 1.0
    tables[:, :, 1] = rng.uniform(0.0, 0.4, (nm, 1)) * np.sin(np.linspace(0, 6, ns))[None, :]
    lesp = rng.uniform(0.05, 0.4, nm)
    for world in (1, 2, 4, 8):
        parts = [uf.ensemble_shard(tables, lesp, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(nm))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        score = np.max(np.abs(tables[:, :, 1]), axis=1) / lesp
        top = set(np.argsort(-score)[:world].tolist())
        assert all(len(top & set(p.tolist())) == 1 for p in parts)
