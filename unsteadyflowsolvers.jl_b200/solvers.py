"""Drop-in `ldvm` / `ldvm2DOF` (+ `ldvmLin`, `lautat`) with the reference's signatures
(src/lowOrder2D/solvers.jl:144, :657, :448, :42): the whole `for istep` loop runs
device-resident inside libunsflow_cuda; this host side only tabulates the prescribed
motion, uploads the state once and downloads it at the end (and at write stamps)."""
import ctypes as C

import numpy as np

from . import _cabi as cabi
from .delvort import delNone, delSpalart
from .kinem import KinemPar, eval_externalvel, eval_kinem
from .typedefs import KinemPar2DOF, TwoDFlowField, TwoDOFPar, TwoDSurf, VortList


def _vorts_struct(v: VortList, keep):
    """unsflow_vorts view of a VortList (arrays kept alive in `keep`)."""
    arrs = [cabi.f64(getattr(v, k)).copy() for k in ("x", "z", "s", "vc", "vx", "vz")]
    keep.append(arrs)
    st = cabi.Vorts()
    st.n = len(v)
    for k, a in zip(("x", "z", "s", "vc", "vx", "vz"), arrs):
        setattr(st, k, cabi.dptr(a) if a.shape[0] else None)
    return st


def surf_struct(surf: TwoDSurf, keep):
    s = cabi.Surf()
    s.ndiv, s.naterm = surf.ndiv, surf.naterm
    s.c, s.uref, s.pvt = surf.c, surf.uref, surf.pvt
    s.lespcrit = float(surf.lespcrit[0])
    s.levflag = int(surf.levflag[0])
    k = surf.kinem
    s.kinem[:] = [k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot]
    s.a0, s.a0dot, s.a0prev = float(surf.a0[0]), float(surf.a0dot[0]), float(surf.a0prev[0])
    for name in ("theta", "x", "cam", "cam_slope", "bnd_x", "bnd_z", "uind", "wind", "downwash", "aterm", "adot", "aprev"):
        a = cabi.f64(getattr(surf, name))
        if a is not getattr(surf, name):
            setattr(surf, name, a)
        setattr(s, name, cabi.dptr(a))
    s.bv = _vorts_struct(surf.bv, keep)
    return s


def field_struct(curfield: TwoDFlowField, keep):
    f = cabi.Field()
    f.u, f.w = float(curfield.u[0]), float(curfield.w[0])
    f.tev = _vorts_struct(curfield.tev, keep)
    f.lev = _vorts_struct(curfield.lev, keep)
    f.extv = _vorts_struct(curfield.extv, keep)
    return f


def kinematics_table(surf: TwoDSurf, curfield: TwoDFlowField, nsteps: int, dt: float, t0: float = 0.0,
                     external=True, two_dof=False):
    """nsteps x 9 table [t, alpha, h, alphadot, hdot, u, udot, U, W] at the running-sum times
    t = t + dt (solvers.jl:180), i.e. what update_externalvel (calcs.jl:75-92) and update_kinem
    (calcs.jl:97-198) would set at each step."""
    tab = np.zeros((nsteps, 9))
    t = t0
    k = KinemPar(**vars(surf.kinem))
    U, W = float(curfield.u[0]), float(curfield.w[0])
    for i in range(nsteps):
        t = t + dt
        if external:
            U, W = eval_externalvel(curfield.velX, curfield.velZ, t, U, W)
        if not two_dof:
            k = eval_kinem(surf.kindef, t, surf.c, surf.uref, k)
        tab[i] = [t, k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot, U, W]
    return tab


class LdvmHandle:
    """RAII wrapper of unsflow_ldvm (unsflow_ldvm_create .. unsflow_ldvm_destroy)."""

    def __init__(self, surf, curfield, opts):
        self._h = C.c_void_p()
        self._keep = []
        self.surf, self.curfield = surf, curfield
        self.surfs = list(surf) if isinstance(surf, (list, tuple)) else [surf]
        self.multi = isinstance(surf, (list, tuple))
        f = field_struct(curfield, self._keep)
        if self.multi:   # ldvm(surf::Vector{TwoDSurf}, ...), solvers.jl:270
            arr = (cabi.Surf * len(self.surfs))(*[surf_struct(q, self._keep) for q in self.surfs])
            cabi.check(cabi.lib().unsflow_ldvm_multi_create(C.byref(self._h), len(self.surfs), arr, C.byref(f), C.byref(opts)))
        else:
            s = surf_struct(surf, self._keep)
            cabi.check(cabi.lib().unsflow_ldvm_create(C.byref(self._h), C.byref(s), C.byref(f), C.byref(opts)))

    def close(self):
        if self._h:
            cabi.lib().unsflow_ldvm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_kinematics(self, table):
        t = cabi.f64(table)
        cabi.check(cabi.lib().unsflow_ldvm_set_kinematics(self._h, t.shape[0], cabi.dptr(t)))

    def set_parallel(self, rank, nranks, nccl_unique_id):
        """unsflow_ldvm_set_parallel: split the roll-up pair tiles over `nranks` GPUs (one process each)"""
        cabi.check(cabi.lib().unsflow_ldvm_set_parallel(self._h, rank, nranks, nccl_unique_id))

    def parallel_info(self):
        """(rank, nranks, peer_exchange): peer_exchange = 1 when the step exchanges through NVLink peer memory"""
        a, b, c = C.c_int(0), C.c_int(1), C.c_int(0)
        cabi.check(cabi.lib().unsflow_ldvm_parallel_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_2dof(self, strpar: TwoDOFPar, kinem: KinemPar2DOF):
        a, b = strpar.as_array(), kinem.as_array()
        cabi.check(cabi.lib().unsflow_ldvm_set_2dof(self._h, cabi.dptr(a), cabi.dptr(b)))

    def get_2dof(self, kinem: KinemPar2DOF):
        b = np.zeros(22)
        cabi.check(cabi.lib().unsflow_ldvm_get_2dof(self._h, cabi.dptr(b)))
        kinem.from_array(b)

    def run(self, nsteps):
        mat = np.zeros((nsteps, 1 + 7 * len(self.surfs) if self.multi else 8), order="F")
        cabi.check(cabi.lib().unsflow_ldvm_run(self._h, nsteps, mat.ctypes.data_as(cabi.c_double_p)))
        return mat

    def timing(self):
        ms, nl = C.c_float(0), C.c_int64(0)
        cabi.check(cabi.lib().unsflow_ldvm_timing(self._h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def counts(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        cabi.check(cabi.lib().unsflow_ldvm_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def download(self):
        """unsflow_ldvm_get_state -> mutate surf and curfield in place, like the reference."""
        fld = self.curfield
        nt, nl, ne = self.counts()
        keep = []
        for v, n in ((fld.tev, nt), (fld.lev, nl), (fld.extv, ne)):
            v.resize(n)
        structs = [surf_struct(q, keep) for q in self.surfs]     # keep: one bv entry per surface, in order
        f = field_struct(fld, keep)                               # then tev, lev, extv
        if self.multi:
            arr = (cabi.Surf * len(structs))(*structs)
            cabi.check(cabi.lib().unsflow_ldvm_multi_get_state(self._h, len(structs), arr, C.byref(f)))
            structs = list(arr)
        else:
            cabi.check(cabi.lib().unsflow_ldvm_get_state(self._h, C.byref(structs[0]), C.byref(f)))
        for surf, s in zip(self.surfs, structs):
            surf.levflag[0] = s.levflag
            k = surf.kinem
            k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot = list(s.kinem)
            surf.a0[0], surf.a0dot[0], surf.a0prev[0] = s.a0, s.a0dot, s.a0prev
        fld.u[0], fld.w[0] = f.u, f.w
        # _vorts_struct made copies: copy back (keep order: bv per surface, tev, lev, extv)
        for v, arrs in zip([q.bv for q in self.surfs] + [fld.tev, fld.lev, fld.extv], keep):
            v.assign(*arrs)


# Shedding-solve mode of the reference-signature entry points when the caller does not choose one:
# the restatement of the reference's own mechanics (NLsolve trust-region iteration on the mutating
# functors, state left at its last finite-difference probe, solvers.jl:199,216 / typedefs.jl:236-247).
# SOLVE_EXACT (state at the exact root) is the opt-in; the two differ by |dCl| ~ 8e-3 on config C1.
DEFAULT_SOLVE_MODE = cabi.SOLVE_NLSOLVE


def _opts(driver, dt, nsteps, delvort, solve_mode):
    if solve_mode is None:
        solve_mode = DEFAULT_SOLVE_MODE
    o = cabi.Opts()
    o.driver, o.solve_mode, o.device = driver, solve_mode, -1
    if isinstance(delvort, delSpalart):
        o.delvort_kind, o.del_limit, o.del_dist, o.del_tol = cabi.DEL_SPALART, delvort.limit, delvort.dist, delvort.tol
    elif delvort is None or isinstance(delvort, delNone):
        o.delvort_kind = cabi.DEL_NONE
    else:
        raise TypeError("delvort must be delNone() or delSpalart(...)")
    o.dt, o.max_steps = dt, nsteps
    return o


def _round_sig(t, nround):
    """directory name "$(round(t, sigdigits=nround))" of solvers.jl:249"""
    return repr(float(f"{t:.{nround}g}"))


def _write_results_summary(mat):
    """solvers.jl:261-264"""
    with open("resultsSummary", "w") as f:
        f.write("#time \talpha(deg) \th/c \tu/uref \tA0 \tCl \tCd \tCm \n")
        for row in mat:
            f.write("\t".join(repr(float(v)) for v in row) + "\n")


def _run(driver, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort, maxwrite, nround,
         solve_mode, strpar=None, kinem=None, write_summary=True):
    if startflag == 0:
        t = 0.0
    elif startflag == 1:
        raise NotImplementedError("startflag=1 (restart from resultsSummary) is broken in the reference "
                                  "(solvers.jl:150-155) and not supported")
    else:
        raise ValueError("invalid start flag, should be 0 or 1")  # solvers.jl:157
    multi = isinstance(surf, (list, tuple))
    two_dof = driver == cabi.DRIVER_LDVM2DOF
    if multi:
        if driver != cabi.DRIVER_LDVM:
            raise TypeError("a Vector{TwoDSurf} is accepted by ldvm only (solvers.jl:270)")
        dt = dtstar * min(q.c / q.uref for q in surf)  # solvers.jl:289
        table = np.ascontiguousarray(np.stack([kinematics_table(q, curfield, nsteps, dt, t) for q in surf], axis=1))
    else:
        dt = dtstar * surf.c / surf.uref  # solvers.jl:161
        table = kinematics_table(surf, curfield, nsteps, dt, t, external=(driver != cabi.DRIVER_LAUTAT), two_dof=two_dof)
    h = LdvmHandle(surf, curfield, _opts(driver, dt, nsteps, delvort, solve_mode))
    try:
        if two_dof:
            h.set_2dof(strpar, kinem)
        h.set_kinematics(table)
        write_steps = []
        if writeflag == 1:  # solvers.jl:164-175
            for i in range(1, maxwrite + 1):
                write_steps.append(int(round(writeInterval * float(i) / dt)))
        mats, done = [], 0
        stops = sorted(set(s for s in write_steps if 0 < s <= nsteps) | {nsteps})
        for stop in stops:
            if stop > done:
                mats.append(h.run(stop - done))
                done = stop
            if stop in write_steps:
                h.download()
                from .postprocess import writeStamp
                tstamp = float(table[stop - 1].reshape(-1)[0])
                if multi:
                    from .postprocess import writeStampMulti
                    writeStampMulti(_round_sig(tstamp, nround), tstamp, surf, curfield)
                else:
                    writeStamp(_round_sig(tstamp, nround), tstamp, surf, curfield)
        h.download()
        if two_dof:
            h.get_2dof(kinem)
        _run.last_timing = h.timing()
    finally:
        h.close()
    mat = np.vstack(mats) if mats else np.zeros((0, 8))
    if multi and write_summary:
        _write_results_summary(mat)
        write_summary = False
    if write_summary:
        _write_results_summary(mat)
    return mat, surf, curfield


def ldvm(surf, curfield, nsteps=500, dtstar=0.015, startflag=0, writeflag=0, writeInterval=1000.0, delvort=None, *,
         maxwrite=50, nround=6, solve_mode=None, write_summary=True):
    """ldvm(surf::TwoDSurf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval, delvort;
    maxwrite, nround) -> (mat, surf, curfield), solvers.jl:144-268.  Passing a list of TwoDSurf
    selects the multi-surface method ldvm(surf::Vector{TwoDSurf}, ...) of solvers.jl:270-445
    (mat then has 1 + 7*nsurf columns)."""
    return _run(cabi.DRIVER_LDVM, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval,
                delvort or delNone(), maxwrite, nround, solve_mode, write_summary=write_summary)


def ldvm2DOF(surf, curfield, strpar, kinem, nsteps=500, dtstar=0.015, startflag=0, writeflag=0, writeInterval=1000.0,
             delvort=None, *, maxwrite=50, nround=6, solve_mode=None, write_summary=True):
    """ldvm2DOF(surf, curfield, strpar::TwoDOFPar, kinem::KinemPar2DOF, ...), solvers.jl:657-788"""
    return _run(cabi.DRIVER_LDVM2DOF, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval,
                delvort or delNone(), maxwrite, nround, solve_mode, strpar=strpar, kinem=kinem,
                write_summary=write_summary)


def ldvmLin(surf, curfield, nsteps=500, dtstar=0.015, startflag=0, writeflag=0, writeInterval=1000.0, delvort=None, *,
            maxwrite=50, nround=6, write_summary=True):
    """ldvmLin(...), solvers.jl:448-655 (closed-form linearised shedding)"""
    return _run(cabi.DRIVER_LDVMLIN, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval,
                delvort or delNone(), maxwrite, nround, cabi.SOLVE_EXACT, write_summary=write_summary)


def lautat(surf, curfield, nsteps=500, dtstar=0.015, startflag=0, writeflag=0, writeInterval=1000.0, delvort=None, *,
           maxwrite=50, nround=6, solve_mode=None, write_summary=True):
    """lautat(...), solvers.jl:42-142 (wakerollup=1)"""
    return _run(cabi.DRIVER_LAUTAT, surf, curfield, nsteps, dtstar, startflag, writeflag, writeInterval,
                delvort or delNone(), maxwrite, nround, solve_mode, write_summary=write_summary)


def ldvm_batch(surf, tables, lespcrit, dtstar=0.015, *, driver=cabi.DRIVER_LDVM, delvort=None, solve_mode=None):
    """Batched ensemble of independent LDVM runs (BASELINE config 3): one thread-block per member
    runs the whole `for istep` loop of solvers.jl:178-257 on the GPU.  All members share the
    geometry / initial state of `surf` (empty wake); member m follows tables[m] (nsteps x 9, see
    kinematics_table) with LESP threshold lespcrit[m].
    Returns (mats [nmember, nsteps, 8], ntev [nmember], nlev [nmember], kernel_ms)."""
    tables = cabi.f64(tables)
    nmember, nsteps, ncol = tables.shape
    if ncol != 9:
        raise ValueError("tables must be nmember x nsteps x 9")
    lesp = cabi.f64(np.broadcast_to(np.asarray(lespcrit, dtype=np.float64), (nmember,)))
    dt = dtstar * surf.c / surf.uref
    keep = []
    sst = surf_struct(surf, keep)
    opts = _opts(driver, dt, nsteps, delvort or delNone(), solve_mode)
    mats = np.zeros((nmember, 8, nsteps))            # member blocks of column-major nsteps x 8
    ntev = np.zeros(nmember, dtype=np.int64)
    nlev = np.zeros(nmember, dtype=np.int64)
    i64p = C.POINTER(C.c_int64)
    cabi.check(cabi.lib().unsflow_ldvm_batch_run(C.byref(sst), C.byref(opts), nmember, nsteps, cabi.dptr(tables), cabi.dptr(lesp),
                                                 mats.ctypes.data_as(cabi.c_double_p), ntev.ctypes.data_as(i64p),
                                                 nlev.ctypes.data_as(i64p)))
    ms = C.c_float(0)
    cabi.check(cabi.lib().unsflow_ldvm_batch_timing(C.byref(ms)))
    return np.ascontiguousarray(mats.transpose(0, 2, 1)), ntev, nlev, ms.value


def ensemble_shard(tables, lespcrit, rank, world, c=1.0):
    """Member indices rank `rank` of `world` should run: members sorted by the same cost heuristic the ensemble
    kernel orders its queues by (largest quasi-steady leading-edge loading the table asks for over the member's LESP
    threshold -- a member that sheds LEVs every step has a wake twice as long) and dealt over the ranks boustrophedon,
    so that every GPU gets the same mix of long and short members.  Members are independent: any split is valid,
    this one only balances the load."""
    tables = np.asarray(tables)
    nmember = tables.shape[0]
    lesp = np.broadcast_to(np.asarray(lespcrit, dtype=np.float64), (nmember,))
    u = np.where(np.abs(tables[:, :, 5]) > 1e-12, np.abs(tables[:, :, 5]), 1.0)
    peak = np.max(np.abs(tables[:, :, 1]) + np.abs(tables[:, :, 3]) * c / u + np.abs(tables[:, :, 4]) / u, axis=1)
    order = np.argsort(-(peak / np.maximum(lesp, 1e-6)), kind="stable")
    pos = np.arange(nmember)
    row, col = pos // world, pos % world
    owner = np.where(row % 2 == 1, world - 1 - col, col)
    return np.sort(order[owner == rank])


def ldvm2DOF_batch(surf, tables, lespcrit, strpars, kinems, dtstar=0.015, *, delvort=None, solve_mode=None):
    """Ensemble of independent ldvm2DOF runs (solvers.jl:657-788), e.g. a sweep over structural
    parameters: member m uses TwoDOFPar strpars[m], starts from KinemPar2DOF kinems[m] and follows
    tables[m] (kinematics_table(..., two_dof=True): only t and the external velocities are read).
    Returns (mats [nmember, nsteps, 8], final KinemPar2DOF list, ntev, nlev, kernel_ms)."""
    tables = cabi.f64(tables)
    nmember, nsteps, ncol = tables.shape
    if ncol != 9 or len(strpars) != nmember or len(kinems) != nmember:
        raise ValueError("tables must be nmember x nsteps x 9 and strpars / kinems one per member")
    lesp = cabi.f64(np.broadcast_to(np.asarray(lespcrit, dtype=np.float64), (nmember,)))
    sp = cabi.f64(np.stack([p.as_array() for p in strpars]))
    k0 = cabi.f64(np.stack([k.as_array() for k in kinems]))
    kout = np.zeros_like(k0)
    dt = dtstar * surf.c / surf.uref
    keep = []
    sst = surf_struct(surf, keep)
    opts = _opts(cabi.DRIVER_LDVM2DOF, dt, nsteps, delvort or delNone(), solve_mode)
    mats = np.zeros((nmember, 8, nsteps))
    ntev = np.zeros(nmember, dtype=np.int64)
    nlev = np.zeros(nmember, dtype=np.int64)
    i64p = C.POINTER(C.c_int64)
    cabi.check(cabi.lib().unsflow_ldvm2dof_batch_run(C.byref(sst), C.byref(opts), nmember, nsteps, cabi.dptr(tables), cabi.dptr(lesp),
                                                     cabi.dptr(sp), cabi.dptr(k0), mats.ctypes.data_as(cabi.c_double_p),
                                                     cabi.dptr(kout), ntev.ctypes.data_as(i64p), nlev.ctypes.data_as(i64p)))
    ms = C.c_float(0)
    cabi.check(cabi.lib().unsflow_ldvm_batch_timing(C.byref(ms)))
    finals = []
    for m in range(nmember):
        k = KinemPar2DOF(0.0, 0.0, 0.0, 0.0, 0.0)
        k.from_array(kout[m])
        finals.append(k)
    return np.ascontiguousarray(mats.transpose(0, 2, 1)), finals, ntev, nlev, ms.value


def lve_table(surf, nsteps, dtstar):
    """Host-tabulated motion for LVE (LVE.jl:57-64): iterations i = 0..nsteps at the running-sum
    times t = -dtstar + dtstar + dtstar ... plus one more row (vor_BFR uses t + dtstar,
    LVE.jl:242).  Columns [t, alpha, h, alphadot, hdot, u, udot, X0, Z0] where alpha, h, ... are
    what update_kinem sets and X0 = -kindef.u(t)*t, Z0 = kindef.h(t) are the frame offsets of
    IFR/BFR (calcs.jl:654-656)."""
    tab = np.zeros((nsteps + 2, 9))
    k = KinemPar(**vars(surf.kinem))
    t = -dtstar
    for i in range(nsteps + 2):
        t += dtstar
        k = eval_kinem(surf.kindef, t, surf.c, surf.uref, k)
        tab[i] = [t, surf.kindef.alpha(t), k.h, k.alphadot, k.hdot, k.u, k.udot, -surf.kindef.u(t) * t, surf.kindef.h(t)]
    return tab


def LVE(surf, curfield, nsteps=500, dtstar=0.015, frameCnt=20, view="square", LEV_Stop=1000, anim=False, animStep=30,
        wflag=False, writeInterval=50):
    """LVE(surf, curfield, nsteps, dtstar, frameCnt, view, LEV_Stop, anim, animStep, wflag, writeInterval)
    -> (frames, IFR_field, mat, TEVhist, LEVhist), src/lowOrder2D/LVE.jl:6-302.  The time loop runs on the
    GPU (unsflow_lve_create + unsflow_ldvm_run); the plotting / animation outputs of the reference
    (`frames`, anim) are not produced (frames = 0, its initial value LVE.jl:21)."""
    from .typedefs import TwoDFlowField as _FF, TwoDSurf as _TS
    if anim or wflag:
        raise NotImplementedError("LVE: anim / wflag outputs are host-side plotting and are not mirrored")
    ifr_surf = _TS(surf.coord_file, surf.pvt, surf.kindef, surf.lespcrit, ndiv=surf.ndiv, camberType="linear")   # LVE.jl:17
    ifr_field = _FF()
    nd, npan = surf.ndiv, surf.ndiv - 1
    # refresh_vortex(surf, vor_loc), LVE.jl:55 (geometry as in LVE.jl:29-41)
    dsv = np.sqrt((surf.x[:-1] - surf.x[1:]) ** 2 + (surf.cam[:-1] - surf.cam[1:]) ** 2)
    csd = np.degrees(np.arcsin((surf.cam[1:] - surf.cam[:-1]) / dsv))
    vor_x = surf.x[:-1] + 0.25 * dsv * np.cos(np.radians(csd))
    vor_z = surf.cam[:-1] + 0.25 * dsv * np.sin(np.radians(csd))
    surf.bv.x[:] = vor_x
    surf.bv.z[:] = vor_z
    table = lve_table(surf, nsteps, dtstar)
    opts = _opts(cabi.DRIVER_LDVM, dtstar, nsteps + 1, delNone(), cabi.SOLVE_EXACT)
    keep = []
    sst = surf_struct(surf, keep)
    fst = field_struct(ifr_field, keep)
    h = C.c_void_p()
    L = cabi.lib()
    cabi.check(L.unsflow_lve_create(C.byref(h), C.byref(sst), C.byref(fst), C.byref(opts), surf.rho, float(LEV_Stop),
                                    float(ifr_surf.x[-1]), float(ifr_surf.cam[-1]), ifr_surf.kinem.u, ifr_surf.kinem.hdot))
    try:
        t9 = cabi.f64(table)
        cabi.check(L.unsflow_ldvm_set_kinematics(h, t9.shape[0], cabi.dptr(t9)))
        niter = nsteps + 1
        m9 = np.zeros((niter, 9), order="F")
        cabi.check(L.unsflow_ldvm_run(h, niter, m9.ctypes.data_as(cabi.c_double_p)))
        nt, nl, ne = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        cabi.check(L.unsflow_ldvm_counts(h, C.byref(nt), C.byref(nl), C.byref(ne)))
        ifr_field.tev.resize(nt.value)
        ifr_field.lev.resize(nl.value)
        keep2 = []
        bnd_keep = (surf.bnd_x.copy(), surf.bnd_z.copy())
        s2 = surf_struct(surf, keep2)
        f2 = field_struct(ifr_field, keep2)
        cabi.check(L.unsflow_ldvm_get_state(h, C.byref(s2), C.byref(f2)))
        circ = np.zeros(nd)
        gcrit = C.c_double(0)
        cabi.check(L.unsflow_lve_get_circ(h, cabi.dptr(circ), C.byref(gcrit)))
        ms, nl_ = C.c_float(0), C.c_int64(0)
        cabi.check(L.unsflow_ldvm_timing(h, C.byref(ms), C.byref(nl_)))
        LVE.last_timing = (ms.value, nl_.value)
    finally:
        L.unsflow_ldvm_destroy(h)
    surf.bnd_x[:], surf.bnd_z[:] = bnd_keep                  # LVE never touches surf.bnd_x/z
    k = surf.kinem
    k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot = list(s2.kinem)
    surf.levflag[0] = s2.levflag
    surf.a0[0] = s2.a0
    for v, arrs in zip((ifr_surf.bv, ifr_field.tev, ifr_field.lev, ifr_field.extv), keep2):
        v.assign(*arrs)
    surf.bv.s[:] = circ[:npan]                               # LVE.jl:156-157
    # curfield = body-frame twin of IFR_field at t + dtstar (LVE.jl:238-242, vor_BFR calcs.jl:732-754)
    al, X0, Z0 = table[-1, 1], table[-1, 7], table[-1, 8]
    pvt = surf.pvt
    for src, dst in ((ifr_field.tev, curfield.tev), (ifr_field.lev, curfield.lev)):
        n = len(src)
        bx = (src.x - X0 - pvt) * np.cos(al) - (src.z - Z0) * np.sin(al) + pvt
        bz = (src.x - X0 - pvt) * np.sin(al) + (src.z - Z0) * np.cos(al)
        dst.assign(bx, bz, src.s.copy(), np.full(n, 0.02 * surf.c), np.zeros(n), np.zeros(n))
    mat = np.ascontiguousarray(m9[1:, :8])                   # the reference records loads for i > 0 only
    tevhist = np.stack([table[:niter, 0], ifr_field.tev.s[:niter]], axis=1)
    shed = m9[:, 8] > 0.5
    levhist = np.stack([table[:niter, 0][shed], ifr_field.lev.s[: int(shed.sum())]], axis=1) if shed.any() else np.zeros((0, 2))
    return 0, ifr_field, mat, tevhist, levhist
