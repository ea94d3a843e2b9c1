"""ctypes wrapper of the CPU oracle (oracle/unsflow_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py; never by the product package.  PARITY UNPINNED (see the
header of unsflow_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libunsflow_oracle.so")
_lib = None
dp = C.POINTER(C.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "unsflow_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        L = C.CDLL(_SO)
        L.o_sim_create.restype = C.c_void_p
        L.o_surf_pack_len.restype = C.c_long
        L.o_sim_count.restype = C.c_long
        L.o_sim_nfev.restype = C.c_long
        L.o_mutual_ind_time.restype = C.c_double
        L.o_ind_vel_slice_mt.restype = C.c_double
        L.o_now.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(dp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ind_vel(sx, sz, ss, svc, tx, tz):
    """ind_vel, calcs.jl:462-477 (reference summation order)"""
    sx, sz, ss, svc, tx, tz = map(_f, (sx, sz, ss, svc, tx, tz))
    u = np.zeros(len(tx))
    w = np.zeros(len(tx))
    lib().o_ind_vel(C.c_long(len(sx)), _p(sx), _p(sz), _p(ss), _p(svc), C.c_long(len(tx)), _p(tx), _p(tz), _p(u), _p(w))
    return u, w


def mutual_ind(x, z, s, vc, vx=None, vz=None):
    """mutual_ind, calcs.jl:491-510 (accumulating)"""
    x, z, s, vc = map(_f, (x, z, s, vc))
    vx = np.zeros(len(x)) if vx is None else _f(vx).copy()
    vz = np.zeros(len(x)) if vz is None else _f(vz).copy()
    lib().o_mutual_ind(C.c_long(len(x)), _p(x), _p(z), _p(s), _p(vc), _p(vx), _p(vz))
    return vx, vz


def ind_vel_truth(sx, sz, ss, svc, tx, tz):
    """long-double per-term + long-double sum; returns (u, w, sum|u terms|, sum|w terms|)"""
    sx, sz, ss, svc, tx, tz = map(_f, (sx, sz, ss, svc, tx, tz))
    n = len(tx)
    u, w, ua, wa = (np.zeros(n) for _ in range(4))
    lib().o_ind_vel_truth(C.c_long(len(sx)), _p(sx), _p(sz), _p(ss), _p(svc), C.c_long(n), _p(tx), _p(tz),
                          _p(u), _p(w), _p(ua), _p(wa))
    return u, w, ua, wa


def mutual_ind_time(x, z, s, vc):
    """times the reference's symmetric loop (single thread); returns (seconds, vx, vz)"""
    x, z, s, vc = map(_f, (x, z, s, vc))
    vx = np.zeros(len(x))
    vz = np.zeros(len(x))
    t = lib().o_mutual_ind_time(C.c_long(len(x)), _p(x), _p(z), _p(s), _p(vc), _p(vx), _p(vz))
    return t, vx, vz


def ind_vel_slice_mt(sx, sz, ss, svc, tx, tz):
    """target-sliced over all host threads; returns (seconds, u, w)"""
    sx, sz, ss, svc, tx, tz = map(_f, (sx, sz, ss, svc, tx, tz))
    u = np.zeros(len(tx))
    w = np.zeros(len(tx))
    t = lib().o_ind_vel_slice_mt(C.c_long(len(sx)), _p(sx), _p(sz), _p(ss), _p(svc), C.c_long(len(tx)), _p(tx), _p(tz),
                                 _p(u), _p(w))
    return t, u, w


def num_threads():
    return int(lib().o_num_threads())


def kinem_eval(kind, params, t):
    v, d = C.c_double(0), C.c_double(0)
    p = _f(params)
    lib().o_kinem_eval(C.c_int(kind), _p(p), C.c_double(t), C.byref(v), C.byref(d))
    return v.value, d.value


SURF_HDR = 14
SURF_ARRAYS = ("theta", "x", "cam", "cam_slope", "bnd_x", "bnd_z", "uind", "wind", "downwash")


def pack_surf(surf):
    """flat surf pack (layout documented above o_surf_pack_len in unsflow_oracle.c) from a
    host-mirror TwoDSurf-like object (duck-typed: tests build it with the product's host
    mirror, which holds no compute)."""
    nd, na = surf.ndiv, surf.naterm
    k = surf.kinem
    hdr = [surf.c, surf.uref, surf.pvt, float(surf.lespcrit[0]), float(surf.levflag[0]),
           k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot, float(surf.a0[0]), float(surf.a0dot[0]), float(surf.a0prev[0])]
    parts = [np.array(hdr, dtype=np.float64)]
    parts += [_f(getattr(surf, n)) for n in SURF_ARRAYS]
    parts += [_f(surf.aterm), _f(surf.adot), _f(surf.aprev)]
    bv = surf.bv
    parts += [_f(bv.x), _f(bv.z), _f(bv.s), _f(bv.vc), _f(bv.vx), _f(bv.vz)]
    p = np.concatenate(parts)
    assert p.shape[0] == SURF_HDR + 9 * nd + 3 * na + 6 * (nd - 1)
    return p


def unpack_surf(p, nd, na):
    out = {"c": p[0], "uref": p[1], "pvt": p[2], "lespcrit": p[3], "levflag": int(p[4]),
           "alpha": p[5], "h": p[6], "alphadot": p[7], "hdot": p[8], "u": p[9], "udot": p[10],
           "a0": p[11], "a0dot": p[12], "a0prev": p[13]}
    q = SURF_HDR
    for n in SURF_ARRAYS:
        out[n] = p[q:q + nd].copy()
        q += nd
    for n in ("aterm", "adot", "aprev"):
        out[n] = p[q:q + na].copy()
        q += na
    nb = nd - 1
    for i, n in enumerate(("bv_x", "bv_z", "bv_s", "bv_vc", "bv_vx", "bv_vz")):
        out[n] = p[q + i * nb:q + (i + 1) * nb].copy()
    return out


DRIVER_LDVM, DRIVER_LDVM2DOF, DRIVER_LDVMLIN, DRIVER_LAUTAT = 0, 1, 2, 3
# shedding-solve mode: 0 = state at the exact root, 1 = restatement of the reference's NLsolve mechanics
# (solvers.jl:199,216; the default, like the product's DEFAULT_SOLVE_MODE)
SOLVER_EXACT, SOLVER_NLSOLVE = 0, 1
DEFAULT_SOLVER_MODE = SOLVER_NLSOLVE


class Sim:
    """Oracle simulation state: TwoDSurf + TwoDFlowField (+ 2DOF) restated in C."""

    def __init__(self, surf, curfield=None):
        self.nd, self.na = surf.ndiv, surf.naterm
        self.h = C.c_void_p(lib().o_sim_create(C.c_int(self.nd), C.c_int(self.na)))
        self.set_surf(surf)
        if curfield is not None:
            self.set_field(curfield)

    def close(self):
        if self.h:
            lib().o_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_surf(self, surf):
        p = pack_surf(surf)
        lib().o_sim_set_surf(self.h, _p(p))

    def set_surf_pack(self, p):
        p = _f(p)
        lib().o_sim_set_surf(self.h, _p(p))

    def get_surf(self):
        n = lib().o_surf_pack_len(C.c_int(self.nd), C.c_int(self.na))
        p = np.zeros(n)
        lib().o_sim_get_surf(self.h, _p(p))
        return unpack_surf(p, self.nd, self.na)

    def set_vorts(self, fam, x, z, s, vc, vx=None, vz=None):
        n = len(x)
        v = np.zeros((6, n))
        for k, a in enumerate((x, z, s, vc, vx, vz)):
            if a is not None:
                v[k] = a
        v = _f(v)
        lib().o_sim_set_vorts(self.h, C.c_int(fam), C.c_long(n), _p(v))

    def set_field(self, fld):
        for fam, v in enumerate((fld.tev, fld.lev, fld.extv)):
            self.set_vorts(fam, v.x, v.z, v.s, v.vc, v.vx, v.vz)
        lib().o_sim_set_field_uw(self.h, C.c_double(float(fld.u[0])), C.c_double(float(fld.w[0])))

    def get_vorts(self, fam):
        n = lib().o_sim_count(self.h, C.c_int(fam))
        v = np.zeros((6, max(n, 1)))
        lib().o_sim_get_vorts(self.h, C.c_int(fam), _p(v))
        v = v.reshape(-1)[: 6 * n].reshape(6, n) if n else np.zeros((6, 0))
        return dict(zip(("x", "z", "s", "vc", "vx", "vz"), v))

    def set_2dof(self, strpar11, kin22):
        a, b = _f(strpar11), _f(kin22)
        lib().o_sim_set_2dof(self.h, _p(a), _p(b))

    def get_2dof(self):
        b = np.zeros(22)
        lib().o_sim_get_2dof(self.h, _p(b))
        return b

    def set_forces(self, cl, cd, cm):
        lib().o_sim_set_forces(self.h, C.c_double(cl), C.c_double(cd), C.c_double(cm))

    def run(self, driver, table, dt, solver_mode=None, delvort=(0, 0, 0.0, 0.0)):
        solver_mode = DEFAULT_SOLVER_MODE if solver_mode is None else solver_mode
        table = _f(table)
        n = table.shape[0]
        mat = np.zeros((n, 8))
        rc = lib().o_sim_run(self.h, C.c_int(driver), C.c_long(n), C.c_double(dt), _p(table), C.c_int(solver_mode),
                             C.c_int(delvort[0]), C.c_long(int(delvort[1])), C.c_double(delvort[2]), C.c_double(delvort[3]),
                             _p(mat))
        assert rc == 0
        return mat

    def wakeroll(self, dt):
        lib().o_sim_wakeroll(self.h, C.c_double(dt))

    def update_indbound(self):
        lib().o_sim_update_indbound(self.h)

    def control_vort_count(self, kind, limit, dist, tol):
        lib().o_sim_control_vort_count(self.h, C.c_int(kind), C.c_long(limit), C.c_double(dist), C.c_double(tol))

    def nfev(self):
        return lib().o_sim_nfev(self.h)

    def delcp_edgevel(self, rho, velx=0.0, velz=0.0):
        """calc_delcp / calc_edgeVel (postprocess.jl:184-244) on the current state -> (p_com, p_in, p_out, q_u, q_l)"""
        out = np.zeros(5 * self.nd)
        lib().o_sim_delcp_edgevel(self.h, C.c_double(rho), C.c_double(velx), C.c_double(velz), _p(out))
        return tuple(out.reshape(5, self.nd))


# synthetic inputs (counter-based generator shared by tests, bench and tools): see /synth.py
import sys as _sys
_sys.path.insert(0, os.path.dirname(_HERE))
from synth import s_wake, splitmix_uniform  # noqa: E402,F401


class MultiSim:
    """Oracle state of the multi-surface ldvm(surf::Vector{TwoDSurf}), solvers.jl:270-445."""

    def __init__(self, surfs, curfield=None):
        self.ns, self.nd, self.na = len(surfs), surfs[0].ndiv, surfs[0].naterm
        assert all(s.ndiv == self.nd and s.naterm == self.na for s in surfs)
        L = lib()
        L.o_msim_create.restype = C.c_void_p
        L.o_msim_count.restype = C.c_long
        L.o_msim_nfev.restype = C.c_long
        self.h = C.c_void_p(L.o_msim_create(C.c_int(self.ns), C.c_int(self.nd), C.c_int(self.na)))
        for i, s in enumerate(surfs):
            p = pack_surf(s)
            L.o_msim_set_surf(self.h, C.c_int(i), _p(p))
        if curfield is not None:
            for fam, v in enumerate((curfield.tev, curfield.lev, curfield.extv)):
                n = len(v)
                a = _f(np.stack([v.x, v.z, v.s, v.vc, v.vx, v.vz])) if n else np.zeros((6, 1))
                L.o_msim_set_vorts(self.h, C.c_int(fam), C.c_long(n), _p(a))
            L.o_msim_set_field_uw(self.h, C.c_double(float(curfield.u[0])), C.c_double(float(curfield.w[0])))

    def close(self):
        if self.h:
            lib().o_msim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, tables, dt, solver_mode=None, delvort=(0, 0, 0.0, 0.0)):
        """tables: nsteps x nsurf x 9"""
        solver_mode = DEFAULT_SOLVER_MODE if solver_mode is None else solver_mode
        tables = _f(tables)
        n = tables.shape[0]
        mat = np.zeros((n, 1 + 7 * self.ns))
        rc = lib().o_msim_run(self.h, C.c_long(n), C.c_double(dt), _p(tables), C.c_int(solver_mode), C.c_int(delvort[0]),
                              C.c_long(int(delvort[1])), C.c_double(delvort[2]), C.c_double(delvort[3]), _p(mat))
        assert rc == 0
        return mat

    def get_surf(self, i):
        n = lib().o_surf_pack_len(C.c_int(self.nd), C.c_int(self.na))
        p = np.zeros(n)
        lib().o_msim_get_surf(self.h, C.c_int(i), _p(p))
        return unpack_surf(p, self.nd, self.na)

    def get_vorts(self, fam):
        n = lib().o_msim_count(self.h, C.c_int(fam))
        v = np.zeros((6, max(n, 1)))
        lib().o_msim_get_vorts(self.h, C.c_int(fam), _p(v))
        v = v.reshape(-1)[: 6 * n].reshape(6, n) if n else np.zeros((6, 0))
        return dict(zip(("x", "z", "s", "vc", "vx", "vz"), v))


class LveSim:
    """Oracle state of LVE (src/lowOrder2D/LVE.jl:6-302)."""

    def __init__(self, surf, ifr_surf, lev_stop=1000.0):
        L = lib()
        L.o_lve_create.restype = C.c_void_p
        L.o_lve_count.restype = C.c_long
        self.nd = surf.ndiv
        p = pack_surf(surf)
        self.h = C.c_void_p(L.o_lve_create(C.c_int(surf.ndiv), C.c_int(surf.naterm), _p(p), C.c_double(surf.rho), C.c_double(lev_stop),
                                           C.c_double(float(ifr_surf.x[-1])), C.c_double(float(ifr_surf.cam[-1])),
                                           C.c_double(ifr_surf.kinem.u), C.c_double(ifr_surf.kinem.hdot)))

    def close(self):
        if self.h:
            lib().o_lve_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, table, nsteps, dtstar):
        table = _f(table)
        assert table.shape == (nsteps + 2, 9)
        mat = np.zeros((nsteps, 8))
        assert lib().o_lve_run(self.h, C.c_long(nsteps), C.c_double(dtstar), _p(table), _p(mat)) == 0
        return mat

    def get_vorts(self, which):
        n = lib().o_lve_count(self.h, C.c_int(which))
        v = np.zeros((6, max(n, 1)))
        lib().o_lve_get_vorts(self.h, C.c_int(which), _p(v))
        v = v.reshape(-1)[: 6 * n].reshape(6, n) if n else np.zeros((6, 0))
        return dict(zip(("x", "z", "s", "vc", "vx", "vz"), v))

    def get_circ(self):
        c = np.zeros(self.nd)
        g = C.c_double(0)
        lib().o_lve_get_circ(self.h, _p(c), C.byref(g))
        return c, g.value
