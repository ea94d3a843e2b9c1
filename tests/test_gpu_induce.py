"""GPU parity tests of the induction kernels through the C ABI (ctypes mirror of the ccall
shim) against the CPU oracle.  Tolerance: <= 1e-12, error normalised by max_i sum_j |term_ij|
against the long-double truth (BASELINE.json north_star / SURVEY.md 8c)."""
import ctypes as C
import math

import numpy as np
import pytest

from cases import c1_surf, rel_err_normalised

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _vl(uf, x, z, s, vc):
    return uf.VortList.from_arrays(x, z, s, vc)


@pytest.mark.parametrize("n,uniform", [(1, True), (7, True), (513, True), (1000, True), (4097, False), (20000, True)])
def test_ind_vel_parity(gpu_lib, uf, oracle, n, uniform):
    x, z, g, vc = oracle.s_wake(n, uniform_vc=uniform)
    nt = min(n, 300)
    tx, tz = x[:nt] + 0.013, z[:nt] - 0.007
    u, w = uf.ind_vel(_vl(uf, x, z, g, vc), tx, tz)
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, tx, tz)
    assert rel_err_normalised(u, tu, au) <= TOL and rel_err_normalised(w, tw, aw) <= TOL
    # and against the reference-order double sum
    ru, rw = oracle.ind_vel(x, z, g, vc, tx, tz)
    assert rel_err_normalised(u, ru, au) <= TOL and rel_err_normalised(w, rw, aw) <= TOL


def test_ind_vel_two_vortex_kat_and_single_source_method(gpu_lib, uf):
    g, vc = 0.7, 0.02
    d = np.array([0.0, 1e-3, 0.02, 0.5, 30.0, 1e3])
    u, w = uf.ind_vel(uf.TwoDVort(0.0, 0.0, g, vc), d, np.zeros_like(d))   # calcs.jl:480-488
    speed = g * d / (2 * math.pi * np.sqrt(vc ** 4 + d ** 4))
    assert np.all(np.abs(u) < 1e-18)
    assert np.all(np.abs(w + speed) <= 4e-16 * np.maximum(speed, 1e-300))


def test_ind_vel_edge_cases(gpu_lib, uf):
    # empty sources -> zeros; empty targets -> empty
    u, w = uf.ind_vel(uf.VortList(), np.array([1.0, 2.0]), np.array([0.0, 0.0]))
    assert np.all(u == 0) and np.all(w == 0)
    u, w = uf.ind_vel(uf.VortList([uf.TwoDVort(0, 0, 1, 0.02)]), np.zeros(0), np.zeros(0))
    assert u.shape == (0,)
    with pytest.raises(ValueError):
        uf.ind_vel(uf.VortList(), np.zeros(2), np.zeros(3))
    # coincident source/target with vc > 0 is an exact zero, like the reference
    u, w = uf.ind_vel(uf.TwoDVort(0.3, -0.2, 5.0, 0.02), np.array([0.3]), np.array([-0.2]))
    assert u[0] == 0.0 and w[0] == 0.0


@pytest.mark.parametrize("n,uniform", [(2, True), (1000, True), (3001, False)])
def test_mutual_ind_parity_and_accumulation(gpu_lib, uf, oracle, n, uniform):
    x, z, g, vc = oracle.s_wake(n, uniform_vc=uniform)
    v = _vl(uf, x, z, g, vc)
    v.vx[:] = 1.5
    v.vz[:] = -0.5
    uf.mutual_ind(v)
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, x, z)
    assert rel_err_normalised(v.vx - 1.5, tu, au) <= TOL and rel_err_normalised(v.vz + 0.5, tw, aw) <= TOL
    rvx, rvz = oracle.mutual_ind(x, z, g, vc)
    assert rel_err_normalised(v.vx - 1.5, rvx, au) <= TOL
    if uniform:   # antisymmetry identity sum_i G_i v_i = 0
        assert abs(np.sum(g * (v.vx - 1.5))) < 1e-12 * np.sum(np.abs(g)) * np.max(au)


def test_wakeroll_and_update_indbound_parity(gpu_lib, uf, oracle):
    surf, fld, nsteps, dtstar = c1_surf(uf)
    x, z, g, vc = oracle.s_wake(1500)
    x = x - x.max() - 0.1          # wake behind the plate (plate sits at x in [-1, 0])
    fld.tev = _vl(uf, x[:1000], z[:1000], g[:1000], vc[:1000])
    fld.lev = _vl(uf, x[1000:1400], z[1000:1400], g[1000:1400], vc[1000:1400])
    fld.extv = _vl(uf, x[1400:], z[1400:], g[1400:], vc[1400:])
    fld.u[0], fld.w[0] = 0.05, -0.02
    rng = np.random.default_rng(3)
    surf.bv.s[:] = 0.01 * rng.standard_normal(69)
    surf.bv.x[:] = 0.5 * (surf.bnd_x[1:] + surf.bnd_x[:-1])
    surf.bv.z[:] = 0.5 * (surf.bnd_z[1:] + surf.bnd_z[:-1])
    sim = oracle.Sim(surf, fld)
    sim.update_indbound()
    sim.wakeroll(dtstar)
    uf.update_indbound(surf, fld)
    uf.wakeroll(surf, fld, dtstar)
    s = sim.get_surf()
    scale = np.max(np.abs(s["uind"])) + np.max(np.abs(s["wind"]))
    assert np.max(np.abs(surf.uind - s["uind"])) <= 1e-12 * scale
    assert np.max(np.abs(surf.wind - s["wind"])) <= 1e-12 * scale
    for fam, v in enumerate((fld.tev, fld.lev, fld.extv)):
        o = sim.get_vorts(fam)
        vs = np.max(np.abs(o["vx"])) + np.max(np.abs(o["vz"]))
        assert np.max(np.abs(v.vx - o["vx"])) <= 1e-12 * vs and np.max(np.abs(v.vz - o["vz"])) <= 1e-12 * vs
        assert np.max(np.abs(v.x - o["x"])) <= 1e-13 and np.max(np.abs(v.z - o["z"])) <= 1e-13


@pytest.mark.parametrize("variant,n", [(1, 3000), (2, 3000), (2, 5121), (0, 70001)])
def test_device_resident_wake_matches_repeated_wakeroll(gpu_lib, uf, oracle, variant, n):
    """unsflow_wake_* (the bench's device-resident path; direct and symmetric pair-tile
    kernels) == the oracle's wakeroll repeated"""
    cabi = uf._cabi
    nbv, nsteps, dt = 69, (3 if n < 10000 else 1), 0.015
    x, z, g, vc = oracle.s_wake(n)
    bx = np.linspace(-1.0, 0.0, nbv); bz = 0.05 * np.ones(nbv); bs = 0.01 * np.cos(np.arange(nbv)); bvc = np.full(nbv, 0.02)
    h = C.c_void_p()
    cabi.check(gpu_lib.unsflow_wake_create(C.byref(h), n, nbv, 0, 1, None, variant))
    try:
        cabi.check(gpu_lib.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc))))
        cabi.check(gpu_lib.unsflow_wake_step(h, nsteps, 0.1, 0.0, dt))
        gx, gz, gvx, gvz = (np.zeros(n) for _ in range(4))
        cabi.check(gpu_lib.unsflow_wake_download(h, *(cabi.dptr(a) for a in (gx, gz, gvx, gvz))))
        ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(gpu_lib.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
        assert nl.value >= 2 * nsteps and ms.value > 0
    finally:
        gpu_lib.unsflow_wake_destroy(h)
    surf, fld, _, _ = c1_surf(uf)
    surf.bv.assign(bx, bz, bs, bvc, np.zeros(nbv), np.zeros(nbv))
    fld.tev = _vl(uf, x, z, g, vc)
    fld.u[0] = 0.1
    sim = oracle.Sim(surf, fld)
    for _ in range(nsteps):
        sim.wakeroll(dt)
    o = sim.get_vorts(0)
    assert np.max(np.abs(gx - o["x"])) < 1e-12 and np.max(np.abs(gz - o["z"])) < 1e-12
    vs = np.max(np.abs(o["vx"]))
    assert np.max(np.abs(gvx - o["vx"])) < 1e-11 * vs


def test_large_n_properties(gpu_lib, uf, oracle):
    """BASELINE-size property checks where the oracle is too slow: antisymmetry identity at
    N = 2e5 and a truth check on a bounded target sample."""
    n = 200_000
    x, z, g, vc = oracle.s_wake(n)
    v = _vl(uf, x, z, g, vc)
    uf.mutual_ind(v)
    idx = np.arange(0, n, n // 64)[:64]
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, x[idx], z[idx])
    assert rel_err_normalised(v.vx[idx], tu, au) <= TOL and rel_err_normalised(v.vz[idx], tw, aw) <= TOL
    assert abs(np.sum(g * v.vx)) <= 1e-12 * np.sum(np.abs(g)) * np.max(au)
    assert abs(np.sum(g * v.vz)) <= 1e-12 * np.sum(np.abs(g)) * np.max(aw)


def test_measurement_helpers(gpu_lib, uf):
    tf, mhz = C.c_double(), C.c_double()
    uf._cabi.check(gpu_lib.unsflow_fp64_peak(C.byref(tf), C.byref(mhz)))
    assert 5.0 < tf.value < 60.0
    e0, e2, e3 = C.c_double(), C.c_double(), C.c_double()
    uf._cabi.check(gpu_lib.unsflow_rsqrt_probe(1 << 22, 1e-8, 1e18, C.byref(e0), C.byref(e2), C.byref(e3)))
    assert e0.value < 1e-5 and e3.value < 1e-15
    print(f"\nFP64 DFMA peak {tf.value:.2f} TFLOP/s (~{mhz.value:.0f} MHz); rsqrt rel err seed {e0.value:.3e} "
          f"newton {e2.value:.3e} cubic {e3.value:.3e}")


def test_fast_precision_mode_stays_inside_the_parity_tolerance(gpu_lib, uf, oracle):
    """unsflow_set_precision(FAST): bias-centred Newton rsqrt, <= 6.5e-13 per pair -- still <= 1e-12
    normalised, but measurably above the round-off level of the default FULL mode"""
    x, z, g, vc = oracle.s_wake(60000)
    idx = np.arange(0, 60000, 1000)
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, x[idx], z[idx])
    errs = {}
    try:
        for mode in (1, 0):
            uf._cabi.check(gpu_lib.unsflow_set_precision(mode))
            v = _vl(uf, x, z, g, vc)
            uf.mutual_ind(v)                       # 60000 >= 16384: symmetric pair-tile kernel
            errs[mode] = max(rel_err_normalised(v.vx[idx], tu, au), rel_err_normalised(v.vz[idx], tw, aw))
    finally:
        gpu_lib.unsflow_set_precision(0)
    assert errs[0] < 1e-14 and 1e-15 < errs[1] <= TOL
    assert gpu_lib.unsflow_set_precision(7) != 0


@pytest.mark.parametrize("ns,nt,uniform", [
    (2047, 255, True), (2048, 256, True), (2049, 257, True),        # sweep-kernel switch (nt <= 256, ns >= 2048)
    (16383, 256, False), (16384, 257, False), (16385, 1023, True),  # 64- vs 512-source tiles of the direct kernel
    (511, 511, True), (512, 512, False), (513, 1, True), (1, 513, True),
])
def test_ind_vel_dispatch_boundaries(gpu_lib, uf, oracle, ns, nt, uniform):
    """sizes on either side of every kernel-selection rule of api.cu / induce.cuh (plan_induce)"""
    x, z, g, vc = oracle.s_wake(ns, uniform_vc=uniform)
    tx, tz, _, _ = oracle.s_wake(nt, seed=77)
    u, w = uf.ind_vel(_vl(uf, x, z, g, vc), tx, tz)
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, tx, tz)
    assert rel_err_normalised(u, tu, au) <= TOL and rel_err_normalised(w, tw, aw) <= TOL


@pytest.mark.parametrize("ns,nt,uniform", [
    (2048, 70, True), (2080, 70, False), (4735, 72, True), (4737, 73, True),     # one chunk per CTA -> several (148 CTAs x 32 sources = 4736)
    (47295, 70, True), (47400, 100, False), (100_001, 128, True), (300_000, 200, True),   # 1 -> 2 CTAs per SM near 4.7e4; RT = 9, 16, 32
])
def test_sweep_chunk_split(gpu_lib, uf, oracle, ns, nt, uniform):
    """k_sweep (sweep.cuh): the wake is cut into chunks of 32 sources dealt evenly over a multiple of the SM count;
    sizes on either side of the chunk / CTA-count rules of sweep_splits (induce.cu), all three target widths"""
    x, z, g, vc = oracle.s_wake(ns, uniform_vc=uniform)
    tx, tz, _, _ = oracle.s_wake(nt, seed=5)
    u, w = uf.ind_vel(_vl(uf, x, z, g, vc), tx, tz)
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, tx, tz)
    assert rel_err_normalised(u, tu, au) <= TOL and rel_err_normalised(w, tw, aw) <= TOL


@pytest.mark.parametrize("n", [2047, 2048, 2049, 4096, 16383, 16384, 16385, 18433, 28672, 32769, 40000, 90_000, 150_000])
def test_mutual_ind_dispatch_boundaries(gpu_lib, uf, oracle, n):
    """direct <-> pair-tile switch (16384), partial last tiles, tile-size multiples; every
    tile edge is sampled"""
    x, z, g, vc = oracle.s_wake(n)
    v = _vl(uf, x, z, g, vc)
    uf.mutual_ind(v)
    edges = np.unique(np.clip(np.concatenate([np.arange(0, n, 1024) - 1, np.arange(0, n, 1024), [n - 2, n - 1]]), 0, n - 1))
    idx = edges if len(edges) <= 96 else edges[np.linspace(0, len(edges) - 1, 96).astype(int)]
    tu, tw, au, aw = oracle.ind_vel_truth(x, z, g, vc, x[idx], z[idx])
    assert rel_err_normalised(v.vx[idx], tu, au) <= TOL and rel_err_normalised(v.vz[idx], tw, aw) <= TOL
    assert abs(np.sum(g * v.vx)) <= 1e-12 * np.sum(np.abs(g)) * np.max(au)
    assert np.all(np.isfinite(v.vx)) and np.all(np.isfinite(v.vz))


def test_meshgrid_velocity_sampling(gpu_lib, uf, oracle):
    """one->many ind_vel over a meshgrid (typedefs.jl:422-512 + LVE.jl:176-183): matrix-shaped
    targets keep their shape; sum over vortices == the reference's per-vortex accumulation"""
    surf, fld, _, _ = c1_surf(uf)
    x, z, g, vc = oracle.s_wake(700)
    tev = _vl(uf, 0.002 * x - 0.9, 0.8 * z, g, vc)
    lev = _vl(uf, 0.001 * x[:50] - 0.5, 0.5 * z[:50] + 0.2, -g[:50], vc[:50])
    m = uf.meshgrid(surf, tev, lev, 0.25, 0.3, 100, "square")
    m.sample_velocity([tev, lev, surf.bv], t=0.3)
    sx, sz, sg, svc = (np.concatenate([getattr(v, k) for v in (tev, lev, surf.bv)]) for k in ("x", "z", "s", "vc"))
    tu, tw, au, aw = oracle.ind_vel_truth(sx, sz, sg, svc, m.x.reshape(-1), m.z.reshape(-1))
    assert m.uMat.shape == (100, 100) and m.t == 0.3
    assert rel_err_normalised((m.uMat - surf.kinem.u).reshape(-1), tu, au) <= TOL
    assert rel_err_normalised((m.wMat - surf.kinem.hdot).reshape(-1), tw, aw) <= TOL
    assert np.allclose(m.velMag, np.hypot(m.uMat, m.wMat))
