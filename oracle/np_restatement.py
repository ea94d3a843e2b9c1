"""Second, independent restatement of the LDVM step in numpy (test infrastructure only).

Purpose: cross-check the C oracle (oracle/unsflow_oracle.c), which follows the reference
literally (mutating functors + iterative root solve), with a restatement that takes a different
algorithmic route: vectorised induction and a CLOSED-FORM solve of the affine Kelvin /
Kelvin-Kutta systems (the route the GPU kernels take).  Agreement of the two on LEV-shedding,
aeroelastic and vortex-merging runs pins the oracle beyond the surveyor's attached-flow probe
(SURVEY.md appendix B).  Formulas: src/lowOrder2D/calcs.jl, typedefs.jl, postprocess.jl,
solvers.jl:144-268 / :657-788 of the reference (cited inline).
"""
import math

import numpy as np


def induce(sx, sz, sg, svc, tx, tz):
    """ind_vel, calcs.jl:462-477 (vectorised; summation order differs from the reference)"""
    if len(sx) == 0 or len(tx) == 0:
        return np.zeros(len(tx)), np.zeros(len(tx))
    xd = sx[None, :] - tx[:, None]
    zd = sz[None, :] - tz[:, None]
    d2 = xd * xd + zd * zd
    den = 2 * math.pi * np.sqrt(svc[None, :] ** 4 + d2 * d2)
    return -(sg[None, :] * zd / den).sum(axis=1), (sg[None, :] * xd / den).sum(axis=1)


def trapz(y, x):
    """simpleTrapz, utils.jl:105-115"""
    return float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1])) / 2.0)


class NpLdvm:
    """State = plain numpy arrays; drivers: ldvm, ldvm2DOF (k2 / sp set), ldvmLin, lautat."""

    def __init__(self, surf, fld):
        s = surf
        self.c, self.uref, self.pvt, self.lespcrit = s.c, s.uref, s.pvt, float(s.lespcrit[0])
        self.nd, self.na = s.ndiv, s.naterm
        self.theta, self.x, self.cam, self.cam_slope = (np.array(a, float) for a in (s.theta, s.x, s.cam, s.cam_slope))
        self.bnd_x, self.bnd_z = np.array(s.bnd_x, float), np.array(s.bnd_z, float)
        self.a0prev, self.aprev = float(s.a0prev[0]), np.array(s.aprev, float)
        self.aterm, self.adot = np.array(s.aterm, float), np.array(s.adot, float)
        self.a0, self.a0dot, self.levflag = float(s.a0[0]), float(s.a0dot[0]), int(s.levflag[0])
        k = s.kinem
        self.kin = [k.alpha, k.h, k.alphadot, k.hdot, k.u, k.udot]
        self.bv = np.array([s.bv.x, s.bv.z, s.bv.s, s.bv.vc], float)
        self.fam = [np.array([v.x, v.z, v.s, v.vc], float).reshape(4, -1) for v in (fld.tev, fld.lev, fld.extv)]
        self.fs = [float(fld.u[0]), float(fld.w[0])]
        self.cl = self.cm = 0.0
        self.k2 = None
        self.sp = None

    def wake(self):
        return np.concatenate(self.fam, axis=1)

    def downwash(self, ui, wi):
        """update_downwash, calcs.jl:49-54"""
        al, _, ald, hd, u, _ = self.kin
        U, W = self.fs
        return (-(u + U) * math.sin(al) - ui * math.sin(al) + (hd - W) * math.cos(al) - wi * math.cos(al)
                - ald * (self.x - self.pvt * self.c)
                + self.cam_slope * (ui * math.cos(al) + (u + U) * math.cos(al) + (hd - W) * math.sin(al) - wi * math.sin(al)))

    def fourier(self, dw, n):
        """A0 (n = 0, calcs.jl:58-60) or A_n (calcs.jl:59-61, :66-72)"""
        if n == 0:
            return -trapz(dw, self.theta) / (self.uref * math.pi)
        return 2.0 * trapz(dw * np.cos(n * self.theta), self.theta) / (self.uref * math.pi)

    def kinem2dof(self, dt):
        """update_kinem2DOF, calcs.jl:604-650"""
        k, p, c, uref = self.k2, self.sp, self.c, self.uref
        k["alpha_pr"], k["h_pr"] = k["alpha"], k["h"]
        for q in ("alphadot", "hdot"):
            k[q + "_pr3"], k[q + "_pr2"], k[q + "_pr"] = k[q + "_pr2"], k[q + "_pr"], k[q]
        for q in ("alphaddot", "hddot"):
            k[q + "_pr3"], k[q + "_pr2"] = k[q + "_pr2"], k[q + "_pr"]
        m11, m12 = 2.0 / c, -p["x_alpha"] * math.cos(k["alpha"])
        m21, m22 = -2.0 * p["x_alpha"] * math.cos(k["alpha"]) / c, p["r_alpha"] ** 2
        R1 = (4 * p["kappa"] * uref ** 2 * self.cl / (math.pi * c * c)
              - 2 * p["w_h"] ** 2 * (p["cubic_h_1"] * k["h"] + p["cubic_h_3"] * k["h"] ** 3) / c
              - p["x_alpha"] * math.sin(k["alpha"]) * k["alphadot"] ** 2)
        R2 = (8 * p["kappa"] * uref ** 2 * self.cm / (math.pi * c * c)
              - p["w_alpha"] ** 2 * p["r_alpha"] ** 2 * (p["cubic_alpha_1"] * k["alpha"] + p["cubic_alpha_3"] * k["alpha"] ** 3))
        det = m11 * m22 - m21 * m12
        k["hddot_pr"] = (m22 * R1 - m12 * R2) / det
        k["alphaddot_pr"] = (-m21 * R1 + m11 * R2) / det
        ab3 = lambda a, b, cc: (dt / 12.0) * (23 * a - 16 * b + 5 * cc)
        k["alphadot"] = k["alphadot_pr"] + ab3(k["alphaddot_pr"], k["alphaddot_pr2"], k["alphaddot_pr3"])
        k["hdot"] = k["hdot_pr"] + ab3(k["hddot_pr"], k["hddot_pr2"], k["hddot_pr3"])
        k["alpha"] = k["alpha_pr"] + ab3(k["alphadot_pr"], k["alphadot_pr2"], k["alphadot_pr3"])
        k["h"] = k["h_pr"] + ab3(k["hdot_pr"], k["hdot_pr2"], k["hdot_pr3"])
        self.kin[0], self.kin[1], self.kin[2], self.kin[3] = k["alpha"], k["h"], k["alphadot"], k["hdot"]

    def spalart(self, fam, limit, D0, V0, lx, lz):
        """one family of controlVortCount, calcs.jl:550-572"""
        v = self.fam[fam]
        gj, xj, zj = v[2, 0], v[0, 0], v[1, 0]
        d_j = math.hypot(xj - lx, zj - lz)
        z_j = math.hypot(xj, zj)
        for i in range(1, min(20, v.shape[1])):
            gk, xk, zk = v[2, i], v[0, i], v[1, i]
            d_k, z_k = math.hypot(xk - lx, zk - lz), math.hypot(xk, zk)
            fact = abs(gj * gk) * abs(z_j - z_k) / (abs(gj + gk) * (D0 + d_j) ** 1.5 * (D0 + d_k) ** 1.5)
            if fact < V0:
                v[0, i] = (abs(gj) * xj + abs(gk) * xk) / abs(gj + gk)
                v[1, i] = (abs(gj) * zj + abs(gk) * zk) / abs(gj + gk)
                v[2, i] += gj
                self.fam[fam] = v[:, 1:].copy()
                return

    def step(self, dt, row, delvort=None, lautat=False):
        """one iteration of solvers.jl:178-257 (ldvm) / :697-778 (ldvm2DOF) / :47-100 (lautat: no update_externalvel,
        no update_indbound -- surf.uind only ever accumulates each step's own new TEV -- and no LEV branch)"""
        two = self.k2 is not None
        t = row[0]
        if not lautat:
            self.fs = [row[7], row[8]]
        if two:
            self.kinem2dof(dt)
        else:
            self.kin = list(row[1:7])
        al, h, ald, hd, u, _ = self.kin
        pc = self.pvt * self.c
        self.bnd_x = self.bnd_x + dt * ((pc - self.x) * math.sin(al) * ald - u + self.cam * math.cos(al) * ald)   # calcs.jl:453-459
        self.bnd_z = self.bnd_z + dt * (hd + (pc - self.x) * math.cos(al) * ald - self.cam * math.sin(al) * ald)
        tev = self.fam[0]
        if tev.shape[1] == 0:                                                                                      # calcs.jl:375-387
            xt, zt = self.bnd_x[-1] + 0.5 * u * dt, self.bnd_z[-1]
        else:
            xt = self.bnd_x[-1] + (tev[0, -1] - self.bnd_x[-1]) / 3.0
            zt = self.bnd_z[-1] + (tev[1, -1] - self.bnd_z[-1]) / 3.0
        vc = 0.02 * self.c
        w = self.wake()
        if lautat:
            if not hasattr(self, "uind_keep"):
                self.uind_keep, self.wind_keep = np.zeros(self.nd), np.zeros(self.nd)
            ub, wb = self.uind_keep, self.wind_keep
        else:
            ub, wb = induce(w[0], w[1], w[2], w[3], self.bnd_x, self.bnd_z)                                        # calcs.jl:35-38
        one = np.ones(1)
        uT, wT = induce(np.array([xt]), np.array([zt]), one, vc * one, self.bnd_x, self.bnd_z)
        kf = self.uref * self.c * math.pi
        lin = lambda uu, ww: (-uu * math.sin(al) - ww * math.cos(al) + self.cam_slope * (uu * math.cos(al) - ww * math.sin(al)))
        dw0, dw1 = self.downwash(ub, wb), lin(uT, wT)
        K = lambda dw: kf * (self.fourier(dw, 0) + self.fourier(dw, 1) / 2.0)
        Kp = kf * (self.a0prev + self.aprev[0] / 2.0)
        # Kelvin (typedefs.jl:249-251) is affine in the TEV strength: K(dw0) + g K_lin(dw1) - Kp + g = 0
        K1 = kf * (-trapz(dw1, self.theta) / (self.uref * math.pi) + trapz(dw1 * np.cos(self.theta), self.theta) / (self.uref * math.pi))
        gt = -(K(dw0) - Kp) / (K1 + 1.0)
        dw = dw0 + gt * dw1
        a0 = self.fourier(dw, 0)
        nrate = self.na if two else 3
        at_pre = np.array([self.fourier(dw, n) for n in range(1, nrate + 1)])
        self.a0dot = (a0 - self.a0prev) / dt                                                                       # calcs.jl:7
        self.adot[:nrate] = (at_pre - self.aprev[:nrate]) / dt
        ui, wi = ub + gt * uT, wb + gt * wT
        gl = None
        if lautat:
            self.uind_keep, self.wind_keep = ui, wi
        if abs(a0) > self.lespcrit and not lautat:                                                                 # solvers.jl:208
            lev = self.fam[1]
            if self.levflag == 0:                                                                                  # calcs.jl:409-426
                xl = self.bnd_x[0] + 0.5 * (u - ald * math.sin(al) * pc + ui[0]) * dt
                zl = self.bnd_z[0] + 0.5 * (-ald * math.cos(al) * pc - hd + wi[0]) * dt
            else:
                xl = self.bnd_x[0] + (lev[0, -1] - self.bnd_x[0]) / 3.0
                zl = self.bnd_z[0] + (lev[1, -1] - self.bnd_z[0]) / 3.0
            uL, wL = induce(np.array([xl]), np.array([zl]), one, vc * one, self.bnd_x, self.bnd_z)
            dw2 = lin(uL, wL)
            f0 = lambda d_: -trapz(d_, self.theta) / (self.uref * math.pi)
            f1 = lambda d_: 2.0 * trapz(d_ * np.cos(self.theta), self.theta) / (self.uref * math.pi)
            lesp = self.lespcrit if a0 > 0 else -self.lespcrit                                                     # typedefs.jl:340-344
            M = np.array([[kf * (f0(dw1) + f1(dw1) / 2) + 1.0, kf * (f0(dw2) + f1(dw2) / 2) + 1.0], [f0(dw1), f0(dw2)]])
            r = np.array([Kp - K(dw0), lesp - f0(dw0) - 0.0])
            gt, gl = np.linalg.solve(M, r)
            dw = dw0 + gt * dw1 + gl * dw2
            ui, wi = ub + gt * uT + gl * uL, wb + gt * wT + gl * wL
            a0 = self.fourier(dw, 0)
            self.fam[1] = np.concatenate([lev, np.array([[xl], [zl], [gl], [vc]])], axis=1)
            self.levflag = 1
        elif not lautat:
            self.levflag = 0
        self.fam[0] = np.concatenate([tev, np.array([[xt], [zt], [gt], [vc]])], axis=1)
        self.aterm = np.array([self.fourier(dw, n) for n in range(1, self.na + 1)])                                # calcs.jl:66-72
        self.a0 = a0
        self.a0prev = a0
        nprev = self.na if two else 3
        self.aprev[:nprev] = self.aterm[:nprev]
        return self._tail(dt, t, ui, wi, delvort)

    def _tail(self, dt, t, ui, wi, delvort):
        """update_bv, controlVortCount, wakeroll, calc_forces and the mat row (common to all drivers)"""
        a0, al, u, hd = self.a0, self.kin[0], self.kin[4], self.kin[3]
        n_ = np.arange(1, self.na + 1)
        gam = (a0 * (1 + np.cos(self.theta)) + (self.aterm[None, :] * np.sin(n_[None, :] * self.theta[:, None])
                                                 * np.sin(self.theta)[:, None]).sum(axis=1)) * self.uref * self.c     # calcs.jl:204-212
        self.bv[2] = (gam[1:] + gam[:-1]) * (self.theta[1] - self.theta[0]) / 2.0
        self.bv[0] = (self.bnd_x[1:] + self.bnd_x[:-1]) / 2.0
        self.bv[1] = (self.bnd_z[1:] + self.bnd_z[:-1]) / 2.0
        if delvort is not None:                                                                                    # calcs.jl:541-602
            limit, D0, V0 = delvort
            mid = int(round(self.nd / 2.0)) - 1
            if self.fam[0].shape[1] > limit:
                self.spalart(0, limit, D0, V0, self.bnd_x[mid], self.bnd_z[mid])
                if self.fam[1].shape[1] > limit:
                    self.spalart(1, limit, D0, V0, self.bnd_x[mid], self.bnd_z[mid])
        w = self.wake()                                                                                            # calcs.jl:222-294
        src = np.concatenate([w, self.bv], axis=1)
        vx, vz = induce(src[0], src[1], src[2], src[3], w[0], w[1])
        vx, vz = vx + self.fs[0], vz + self.fs[1]
        o = 0
        for f in range(3):
            m = self.fam[f].shape[1]
            self.fam[f][0] += dt * vx[o:o + m]
            self.fam[f][1] += dt * vz[o:o + m]
            o += m
        U, W = self.fs                                                                                             # postprocess.jl:1-32
        uf_ = (u + U) * math.cos(al) / self.uref + (hd - W) * math.sin(al) / self.uref
        cnc = 2 * math.pi * uf_ * (a0 + self.aterm[0] / 2.0)
        cnnc = 2 * math.pi * (3 * self.c * self.a0dot / (4 * self.uref) + self.c * self.adot[0] / (4 * self.uref) + self.c * self.adot[1] / (8 * self.uref))
        cs = 2 * math.pi * a0 * a0
        vv = (ui[:-1] * math.cos(al) - wi[:-1] * math.sin(al)) * self.bv[2]
        nonl = vv.sum() * 2.0 / (self.uref ** 2 * self.c)
        nonl_m = (vv * self.x[:-1]).sum() * 2.0 / (self.uref ** 2 * self.c ** 2)
        cn = cnc + cnnc + nonl
        cl, cd = cn * math.cos(al) + cs * math.sin(al), cn * math.sin(al) - cs * math.cos(al)
        cm = cn * self.pvt - 2 * math.pi * (uf_ * (a0 / 4.0 + self.aterm[0] / 4.0 - self.aterm[1] / 8.0)
                                            + (self.c / self.uref) * (7.0 * self.a0dot / 16.0 + 3.0 * self.adot[0] / 16.0
                                                                      + self.adot[1] / 16.0 - self.adot[2] / 64.0)) - nonl_m
        self.cl, self.cm = cl, cm
        return [t, al, self.kin[1], u, a0, cl, cd, cm]

    def step_lin(self, dt, row, delvort=None):
        """one iteration of ldvmLin, solvers.jl:488-644: linearised closed-form shedding through the integrals
        I1, J1 (downwash without this step's vortices), I2, J2 (unit TEV), I3, J3 (unit LEV)"""
        t = row[0]
        self.fs = [row[7], row[8]]
        self.kin = list(row[1:7])
        al, h, ald, hd, u, _ = self.kin
        pc = self.pvt * self.c
        self.bnd_x = self.bnd_x + dt * ((pc - self.x) * math.sin(al) * ald - u + self.cam * math.cos(al) * ald)
        self.bnd_z = self.bnd_z + dt * (hd + (pc - self.x) * math.cos(al) * ald - self.cam * math.sin(al) * ald)
        w = self.wake()
        ui, wi = induce(w[0], w[1], w[2], w[3], self.bnd_x, self.bnd_z)                                            # :502
        T1 = self.downwash(ui, wi)                                                                                 # :505-510
        cm1 = np.cos(self.theta) - 1.0
        I1 = self.c * trapz(T1 * cm1, self.theta)
        J1 = -trapz(T1, self.theta) / (self.uref * math.pi)
        tev, vc = self.fam[0], 0.02 * self.c
        if tev.shape[1] == 0:                                                                                      # :515-523
            xt, zt = self.bnd_x[-1] + 0.5 * u * dt, self.bnd_z[-1]
        else:
            xt = self.bnd_x[-1] + (tev[0, -1] - self.bnd_x[-1]) / 3.0
            zt = self.bnd_z[-1] + (tev[1, -1] - self.bnd_z[-1]) / 3.0

        def unit(xv, zv):                                                                                          # :525-530
            xd, zd = self.bnd_x - xv, self.bnd_z - zv
            dsq = xd * xd + zd * zd
            return (self.cam_slope * zd + xd) / (2 * math.pi * np.sqrt(dsq ** 2 + vc ** 4))
        T2 = unit(xt, zt)
        sig_prev = -self.uref * self.c * math.pi * (self.a0prev + self.aprev[0] / 2.0)                             # :533
        I2 = trapz(T2 * cm1, self.theta)
        J2 = -trapz(T2, self.theta) / (math.pi * self.uref)
        gt = -(I1 + sig_prev) / (1 + I2)                                                                           # :538
        a0 = J1 + J2 * gt
        c1 = np.array([trapz(T1 * np.cos(n * self.theta), self.theta) for n in range(1, self.na + 1)])
        c2 = np.array([trapz(T2 * np.cos(n * self.theta), self.theta) for n in range(1, self.na + 1)])
        self.aterm = 2.0 * (c1 + gt * c2) / (math.pi * self.uref)                                                  # :542-544
        self.a0dot = (a0 - self.a0prev) / dt                                                                       # :547-550
        self.adot = (self.aterm - self.aprev) / dt
        if abs(a0) > self.lespcrit:                                                                                # :553
            lesp = self.lespcrit if a0 >= 0.0 else -self.lespcrit
            lev = self.fam[1]
            if self.levflag == 0:
                xl = self.bnd_x[0] + 0.5 * (u - ald * math.sin(al) * pc + ui[0]) * dt
                zl = self.bnd_z[0] + 0.5 * (-ald * math.cos(al) * pc - hd + wi[0]) * dt
            else:
                xl = self.bnd_x[0] + (lev[0, -1] - self.bnd_x[0]) / 3.0
                zl = self.bnd_z[0] + (lev[1, -1] - self.bnd_z[0]) / 3.0
            T3 = unit(xl, zl)
            I3 = trapz(T3 * cm1, self.theta)
            J3 = -trapz(T3, self.theta) / (math.pi * self.uref)
            det = J3 * (I2 + 1.0) - J2 * (I3 + 1.0)                                                                # :581-584
            gt = (-J3 * (I1 + sig_prev) + (I3 + 1) * (J1 - lesp)) / det
            gl = (J2 * (I1 + sig_prev) - (I2 + 1) * (J1 - lesp)) / det
            a0 = J1 + J2 * gt + J3 * gl
            c3 = np.array([trapz(T3 * np.cos(n * self.theta), self.theta) for n in range(1, self.na + 1)])
            self.aterm = 2.0 * (c1 + gt * c2 + gl * c3) / (math.pi * self.uref)
            self.fam[1] = np.concatenate([lev, np.array([[xl], [zl], [gl], [vc]])], axis=1)
            self.levflag = 1
        else:
            self.levflag = 0
        self.fam[0] = np.concatenate([tev, np.array([[xt], [zt], [gt], [vc]])], axis=1)
        self.a0 = a0
        self.a0prev = a0                                                                                           # :607-610
        self.aprev[:3] = self.aterm[:3]
        return self._tail(dt, t, ui, wi, delvort)

    def run(self, table, dt, delvort=None, driver="ldvm"):
        if driver == "ldvmLin":
            return np.array([self.step_lin(dt, row, delvort) for row in table])
        return np.array([self.step(dt, row, delvort, lautat=(driver == "lautat")) for row in table])
