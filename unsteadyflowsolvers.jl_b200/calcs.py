"""Drop-ins for the reference's fine-grained hot-path functions (src/lowOrder2D/calcs.jl),
each a thin call into libunsflow_cuda through the C ABI (include/unsflow.h)."""
import numpy as np

from . import _cabi as cabi
from .typedefs import TwoDVort, VortList


def _as_vortlist(src):
    if isinstance(src, VortList):
        return src
    if isinstance(src, TwoDVort):
        return VortList([src])
    return VortList(list(src))


def ind_vel(src, t_x, t_z):
    """ind_vel(src, t_x, t_z) -> (uind, wind), calcs.jl:462-477 (Vector{TwoDVort}) and
    :480-488 (single TwoDVort)."""
    v = _as_vortlist(src)
    tx, tz = cabi.f64(t_x), cabi.f64(t_z)
    if tx.shape != tz.shape:
        raise ValueError("t_x and t_z must have the same length")
    shape = tx.shape                      # matrix-shaped targets (a meshgrid, calcs.jl:480-488 broadcasts) keep their shape
    tx, tz = tx.reshape(-1), tz.reshape(-1)
    u = np.zeros(tx.shape[0])
    w = np.zeros(tx.shape[0])
    sx, sz, ss, svc = (cabi.f64(a) for a in (v.x, v.z, v.s, v.vc))
    cabi.check(cabi.lib().unsflow_ind_vel(len(v), cabi.dptr(sx), cabi.dptr(sz), cabi.dptr(ss), cabi.dptr(svc),
                                          tx.shape[0], cabi.dptr(tx), cabi.dptr(tz), cabi.dptr(u), cabi.dptr(w)))
    return u.reshape(shape), w.reshape(shape)


def mutual_ind(vorts):
    """mutual_ind(vorts) -> vorts, calcs.jl:491-510: accumulates into each vortex's vx, vz."""
    v = _as_vortlist(vorts)
    x, z, s, vc, vx, vz = (cabi.f64(a) for a in (v.x, v.z, v.s, v.vc, v.vx, v.vz))
    cabi.check(cabi.lib().unsflow_mutual_ind(len(v), cabi.dptr(x), cabi.dptr(z), cabi.dptr(s), cabi.dptr(vc),
                                             cabi.dptr(vx), cabi.dptr(vz)))
    v.vx[:] = vx
    v.vz[:] = vz
    if v is not vorts and not isinstance(vorts, TwoDVort):
        for dst, src in zip(vorts, v):
            dst.vx, dst.vz = src.vx, src.vz
    return vorts


def update_indbound(surf, curfield):
    """update_indbound(surf, curfield) -> surf, calcs.jl:35-38"""
    fams = (curfield.tev, curfield.lev, curfield.extv)
    cat = [np.concatenate([cabi.f64(getattr(f, k)) for f in fams]) for k in ("x", "z", "s", "vc")]
    n = cat[0].shape[0]
    u = np.zeros(surf.ndiv)
    w = np.zeros(surf.ndiv)
    bx, bz = cabi.f64(surf.bnd_x), cabi.f64(surf.bnd_z)
    cabi.check(cabi.lib().unsflow_update_indbound(n, *(cabi.dptr(a) for a in cat), surf.ndiv,
                                                  cabi.dptr(bx), cabi.dptr(bz), cabi.dptr(u), cabi.dptr(w)))
    surf.uind[: surf.ndiv] = u
    surf.wind[: surf.ndiv] = w
    return surf


def wakeroll(surf, curfield, dt):
    """wakeroll(surf::TwoDSurf, curfield, dt) -> curfield, calcs.jl:222-294"""
    fams = (curfield.tev, curfield.lev, curfield.extv)
    x, z, s, vc = (np.concatenate([cabi.f64(getattr(f, k)) for f in fams]) for k in ("x", "z", "s", "vc"))
    n = x.shape[0]
    vx = np.zeros(n)
    vz = np.zeros(n)
    bv = surf.bv
    bx, bz, bs, bvc = (cabi.f64(a) for a in (bv.x, bv.z, bv.s, bv.vc))
    cabi.check(cabi.lib().unsflow_wakeroll(n, cabi.dptr(x), cabi.dptr(z), cabi.dptr(s), cabi.dptr(vc),
                                           cabi.dptr(vx), cabi.dptr(vz), len(bv), cabi.dptr(bx), cabi.dptr(bz),
                                           cabi.dptr(bs), cabi.dptr(bvc), float(curfield.u[0]), float(curfield.w[0]),
                                           float(dt)))
    o = 0
    for f in fams:
        m = len(f)
        f.x[:] = x[o:o + m]
        f.z[:] = z[o:o + m]
        f.vx[:] = vx[o:o + m]
        f.vz[:] = vz[o:o + m]
        o += m
    return curfield
