"""Host-side mirror of the snapshot writers (src/lowOrder2D/postprocess.jl:34-244).

Runs on the HOST from a state downloaded with unsflow_ldvm_get_state at the scheduled write
steps (solvers.jl:164-175, :247-252); in the Julia drop-in the reference's own `writeStamp` is
called instead (julia/UnsflowCUDA.jl).  Output-only code, not on the GPU path; the on-disk
layout `Step Files/<t>/{timeKinem,FourierCoeffs,tev,lev,boundv,delcp_edgevel}` is what
`makeVortPlots2D` and the UI read (plots2D.jl:32-175).
"""
import math
import os
import shutil

import numpy as np


def _writedlm(f, mat):
    mat = np.atleast_2d(np.asarray(mat, dtype=np.float64))
    for row in mat:
        f.write("\t".join(repr(float(v)) for v in row) + "\n")


def calc_delcp(surf, vels):
    """calc_delcp, postprocess.jl:184-222 -> (p_com, p_in, p_out)"""
    nd = surf.ndiv
    k = surf.kinem
    a0, a0dot = float(surf.a0[0]), float(surf.a0dot[0])
    p_in, p_out, gam, gamint, p_com, gammod = (np.zeros(nd) for _ in range(6))
    ca, sa = math.cos(k.alpha), math.sin(k.alpha)
    for i in range(nd):
        udash = k.u * ca + k.hdot * sa + surf.uind[i] * ca - surf.wind[i] * sa
        p_in[i] = (8 * surf.uref * math.sqrt(surf.c) * a0 * math.sqrt(surf.x[i]) * udash / (surf.rho * surf.c + 2 * surf.x[i])
                   + 4 * surf.uref * math.sqrt(surf.c) * a0dot * math.sqrt(surf.x[i]))
    for i in range(1, nd):
        th = surf.theta[i]
        udash = k.u * ca + k.hdot * sa + surf.uind[i] * ca - surf.wind[i] * sa
        gam[i] = a0 / math.tan(th / 2)
        gamint[i] = a0dot * (th + math.sin(th))
        for n in range(1, surf.naterm + 1):
            gam[i] += surf.aterm[n - 1] * math.sin(n * th)
            gammod[i] += surf.aterm[n - 1] * math.sin(n * th)
            if n == 1:
                gamint[i] += surf.adot[0] * (th / 2 - math.sin(2 * th) / 4)
            else:
                gamint[i] += surf.adot[n - 1] / 2 * (math.sin((n - 1) * th) / (n - 1) - math.sin((n + 1) * th) / (n + 1))
        gam[i] *= surf.uref
        gammod[i] *= surf.uref
        gamint[i] *= surf.uref * surf.c
        p_out[i] = 2 * udash * gam[i] + gamint[i]
        p_com[i] = (2 * udash * surf.uref * a0 * (math.cos(th / 2) - 1) / math.sin(th / 2) + 2 * udash * gammod[i] + gamint[i]
                    + 8 * surf.uref * udash * surf.c * a0 * math.sin(th / 2) / (surf.rho * surf.c + 2 * surf.c * (math.sin(th / 2)) ** 2))
    return p_com, p_in, p_out


def calc_edgeVel(surf, vels):
    """calc_edgeVel, postprocess.jl:224-244 -> (q_com_u, q_com_l)"""
    nd = surf.ndiv
    k = surf.kinem
    a0 = float(surf.a0[0])
    gammod, qu, ql = (np.zeros(nd) for _ in range(3))
    qu[0] = surf.uref * a0 / math.sqrt(0.5 * surf.rho)
    ql[0] = surf.uref * a0 / math.sqrt(0.5 * surf.rho)
    ca, sa = math.cos(k.alpha), math.sin(k.alpha)
    for i in range(1, nd):
        th = surf.theta[i]
        udash = (k.u + vels[0]) * ca + (k.hdot - vels[1]) * sa + surf.uind[i] * ca - surf.wind[i] * sa
        for n in range(1, surf.naterm + 1):
            gammod[i] += surf.aterm[n - 1] * math.sin(n * th)
        gammod[i] *= surf.uref
        t1 = surf.uref * a0 * (math.cos(th / 2) - 1) / math.sin(th / 2)
        den = math.sqrt(surf.x[i] + surf.rho * surf.c / 2)
        qu[i] = t1 + surf.uref * gammod[i] + (math.sqrt(surf.c) * surf.uref * a0 + math.sqrt(surf.x[i]) * udash) / den
        ql[i] = -t1 - surf.uref * gammod[i] + (-math.sqrt(surf.c) * surf.uref * a0 + math.sqrt(surf.x[i]) * udash) / den
    return qu, ql


def writeStamp(dirname, t, surf, curfield):
    """writeStamp(dirname, t, surf::TwoDSurf, curfield), postprocess.jl:34-113"""
    root = "Step Files"
    os.makedirs(root, exist_ok=True)
    d = os.path.join(root, dirname)
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.mkdir(d)
    k = surf.kinem
    with open(os.path.join(d, "timeKinem"), "w") as f:
        f.write("#time \talpha (deg) \th/c \tu/uref \t\n")
        _writedlm(f, [[t, k.alpha, k.h, k.u]])
    with open(os.path.join(d, "FourierCoeffs"), "w") as f:
        f.write("#Fourier coeffs (0-n) \td/dt (Fourier coeffs) \n")
        m = np.zeros((surf.naterm + 1, 2))
        m[:, 0] = np.concatenate([[surf.a0[0]], surf.aterm])
        m[:, 1] = np.concatenate([[surf.a0dot[0]], surf.adot])
        _writedlm(f, m)
    for name, v in (("tev", curfield.tev), ("lev", curfield.lev), ("boundv", surf.bv)):
        with open(os.path.join(d, name), "w") as f:
            f.write("#strength \tx-position \tz-position  \n")
            if len(v):
                _writedlm(f, np.stack([v.s, v.x, v.z], axis=1))
    with open(os.path.join(d, "delcp_edgevel"), "w") as f:
        f.write("#x \tdelcp \tdelcp_inner \tdelcp_outer \tqu \tql  \n")
        vels = [float(curfield.u[0]), float(curfield.w[0])]
        delcp, delcp_in, delcp_out = calc_delcp(surf, vels)
        qu, ql = calc_edgeVel(surf, vels)
        _writedlm(f, np.stack([surf.x, delcp, delcp_in, delcp_out, qu, ql], axis=1))


def writeStampMulti(dirname, t, surfs, curfield):
    """writeStamp(dirname, t, surf::Vector{TwoDSurf}, curfield), postprocess.jl:115-182.  Unlike the
    single-surface method it writes into `<cwd>/<dirname>` (no "Step Files" level, :116-121), one
    timeKinem-i / FourierCoeffs-i / boundv-i per surface, and skips the zero-core placeholder LEVs
    (`vc != 0`, :149-155)."""
    if os.path.isdir(dirname):
        shutil.rmtree(dirname)
    os.mkdir(dirname)
    for i, surf in enumerate(surfs, start=1):
        k = surf.kinem
        with open(os.path.join(dirname, f"timeKinem-{i}"), "w") as f:
            f.write("#time \talpha (deg) \th/c \tu/uref \t\n")
            _writedlm(f, [[t, k.alpha, k.h, k.u]])
        with open(os.path.join(dirname, f"FourierCoeffs-{i}"), "w") as f:
            f.write("#Fourier coeffs (0-n) \td/dt (Fourier coeffs)  \n")
            m = np.zeros((surf.naterm + 1, 2))
            m[:, 0] = np.concatenate([[surf.a0[0]], surf.aterm])
            m[:, 1] = np.concatenate([[surf.a0dot[0]], surf.adot])
            _writedlm(f, m)
        with open(os.path.join(dirname, f"boundv-{i}"), "w") as f:
            f.write("#strength \tx-position \tz-position  \n")
            _writedlm(f, np.stack([surf.bv.s, surf.bv.x, surf.bv.z], axis=1))
    with open(os.path.join(dirname, "tev"), "w") as f:
        f.write("#strength \tx-position \tz-position  \n")
        if len(curfield.tev):
            _writedlm(f, np.stack([curfield.tev.s, curfield.tev.x, curfield.tev.z], axis=1))
    with open(os.path.join(dirname, "lev"), "w") as f:
        f.write("#strength \tx-position \tz-position  \n")
        lv = curfield.lev
        if len(lv):
            keep = lv.vc != 0.0
            if keep.any():
                _writedlm(f, np.stack([lv.s[keep], lv.x[keep], lv.z[keep]], axis=1))


class meshgrid:
    """meshgrid(surf, tevs, levs, offset, t, width=100, view="square"), typedefs.jl:422-512: the
    sampling window for flow visualisation around the surface in the inertial frame -- mesh
    coordinates `x, z` (rows top to bottom), `uMat = kinem.u`, `wMat = kinem.hdot`, `velMag = 0`,
    the camber line in the inertial frame and the TEV / LEV positions inside the window.
    `sample_velocity(vorts)` adds the velocity the vortices induce at every mesh point (the
    disabled block of LVE.jl:166-190: a one->many `ind_vel` per vortex; here ONE device call over
    all sources through unsflow_ind_vel)."""

    def __init__(self, surf, tevs, levs, offset, t, width=100, view="square"):
        far = surf.x[-1] + surf.c * offset
        near = surf.x[0] - surf.c * offset
        zb = (far - near) / 2
        if view == "wake":
            far = 2 * far
        elif view == "largewake":
            far, zb = 3 * far, 2 * zb
        elif view == "longwake":
            far, zb = 5 * far, 2 * zb
        elif view == "UI Window":
            far = 2 * far
            zb = 3 / 4 * far
        X0 = -surf.kindef.u(t) * t
        Z0 = surf.kindef.h(t)
        lowX, uppX = near + X0 - surf.pvt * surf.c, far + X0 - surf.pvt * surf.c
        lowZ, uppZ = -zb + Z0, zb + Z0
        alpha = surf.kindef.alpha(t)                                   # IFR, calcs.jl:652-663
        self.camX = (surf.x - surf.pvt) * math.cos(alpha) + surf.cam * math.sin(alpha) + X0 + surf.pvt
        self.camZ = -(surf.x - surf.pvt) * math.sin(alpha) + surf.cam * math.cos(alpha) + Z0
        step = (far - near) / (width - 1)
        xs = lowX + step * np.arange(int(math.floor((uppX - lowX) / step + 1e-9)) + 1)     # lowX:step:uppX
        # `height = zBnd*2/step + 1` rows, each a copy of the x range (typedefs.jl:466-467; Julia
        # truncates the comprehension range 1:height to floor(height) rows)
        nrow_x = int(math.floor(zb * 2 / step + 1 + 1e-9))
        zs = uppZ - step * np.arange(int(math.floor((uppZ - lowZ) / step + 1e-9)) + 1)     # uppZ:-step:lowZ
        self.x = np.tile(xs, (nrow_x, 1))
        self.z = np.tile(zs[:, None], (1, xs.shape[0]))
        self.uMat = np.zeros_like(self.x) + surf.kinem.u
        self.wMat = np.zeros_like(self.x) + surf.kinem.hdot
        self.velMag = np.zeros_like(self.x)
        self.t = 0.0
        self.alpha = surf.kinem.alpha

        def inside(v):
            vx, vz = np.asarray(v.x, float), np.asarray(v.z, float)
            m = (vx >= lowX) & (vx <= uppX) & (vz >= lowZ) & (vz <= uppZ)
            return np.stack([vx[m], vz[m]], axis=1)
        self.tev = inside(tevs)
        self.lev = inside(levs)

    def sample_velocity(self, vorts, t=None):
        """uMat, wMat += velocity induced by `vorts` (VortList or list of them); velMag updated."""
        from .calcs import ind_vel
        from .typedefs import VortList
        if isinstance(vorts, (list, tuple)):
            parts = [v for v in vorts if len(v)]
            vorts = VortList.from_arrays(*(np.concatenate([np.asarray(getattr(v, k), float) for v in parts])
                                           for k in ("x", "z", "s", "vc"))) if parts else VortList()
        if self.x.shape != self.z.shape:
            raise ValueError("mesh x and z differ in shape (non-square view): sample on matching sub-blocks")
        u, w = ind_vel(vorts, self.x, self.z)
        self.uMat = self.uMat + u
        self.wMat = self.wMat + w
        self.velMag = np.sqrt(self.uMat ** 2 + self.wMat ** 2)
        if t is not None:
            self.t = t
        return self
