"""Host-side mirror of the reference's data types (src/lowOrder2D/typedefs.jl).

The reference stores Vector{TwoDVort} as an array of boxed 48-byte structs (typedefs.jl:11-18).
Here a vortex family is a `VortList`: SoA float64 arrays (x, z, s, vc, vx, vz) -- the layout
that is uploaded to HBM as is -- with list-like access that hands out `TwoDVort` views.
"""
import math
from dataclasses import dataclass, field as dc_field

import numpy as np

from .kinem import (ConstDef, CosDef, EldRampReturnDef, EldUpDef, EldUpIntDef, FileDef, KinemDef, KinemPar, MotionDef, SinDef)

_FIELDS = ("x", "z", "s", "vc", "vx", "vz")


class TwoDVort:
    """TwoDVort(x, z, s, vc, vx, vz), typedefs.jl:11-18.  Either a free-standing record or a
    view of element `i` of a VortList (writes go through)."""

    __slots__ = ("_owner", "_i", "_v")

    def __init__(self, x=0.0, z=0.0, s=0.0, vc=0.0, vx=0.0, vz=0.0, _owner=None, _i=0):
        object.__setattr__(self, "_owner", _owner)
        object.__setattr__(self, "_i", _i)
        object.__setattr__(self, "_v", None if _owner is not None else [float(x), float(z), float(s), float(vc), float(vx), float(vz)])

    def __getattr__(self, k):
        if k in _FIELDS:
            if self._owner is not None:
                return float(getattr(self._owner, k)[self._i])
            return self._v[_FIELDS.index(k)]
        raise AttributeError(k)

    def __setattr__(self, k, val):
        if k in _FIELDS:
            if self._owner is not None:
                getattr(self._owner, k)[self._i] = val
            else:
                self._v[_FIELDS.index(k)] = float(val)
        else:
            raise AttributeError(k)

    def __repr__(self):
        return "TwoDVort(" + ", ".join(f"{getattr(self, k)!r}" for k in _FIELDS) + ")"


class VortList:
    """Vector{TwoDVort} as structure-of-arrays with amortised push! and popfirst!."""

    def __init__(self, items=()):
        self._cap = 64
        self._n = 0
        self._buf = np.zeros((6, self._cap), dtype=np.float64)
        for it in items:
            self.append(it)

    @classmethod
    def from_arrays(cls, x, z, s, vc, vx=None, vz=None):
        o = cls()
        n = len(x)
        o._cap = max(64, n)
        o._buf = np.zeros((6, o._cap), dtype=np.float64)
        o._n = n
        for k, a in enumerate((x, z, s, vc, vx, vz)):
            if a is not None:
                o._buf[k, :n] = a
        return o

    def __len__(self):
        return self._n

    def _arr(self, k):
        return self._buf[k, : self._n]

    x = property(lambda self: self._arr(0))
    z = property(lambda self: self._arr(1))
    s = property(lambda self: self._arr(2))
    vc = property(lambda self: self._arr(3))
    vx = property(lambda self: self._arr(4))
    vz = property(lambda self: self._arr(5))

    def __getitem__(self, i):
        if isinstance(i, slice):
            idx = range(*i.indices(self._n))
            return [TwoDVort(_owner=self, _i=j) for j in idx]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return TwoDVort(_owner=self, _i=i)

    def __iter__(self):
        for i in range(self._n):
            yield TwoDVort(_owner=self, _i=i)

    def append(self, v):  # push!
        if self._n == self._cap:
            self._cap *= 2
            nb = np.zeros((6, self._cap), dtype=np.float64)
            nb[:, : self._n] = self._buf[:, : self._n]
            self._buf = nb
        self._buf[:, self._n] = [v.x, v.z, v.s, v.vc, v.vx, v.vz]
        self._n += 1

    push = append

    def popfirst(self):  # popfirst!
        v = TwoDVort(*self._buf[:, 0])
        self._buf[:, : self._n - 1] = self._buf[:, 1 : self._n].copy()
        self._n -= 1
        return v

    def resize(self, n):
        if n > self._cap:
            nb = np.zeros((6, n), dtype=np.float64)
            nb[:, : self._n] = self._buf[:, : self._n]
            self._buf, self._cap = nb, n
        self._n = n

    def assign(self, x, z, s, vc, vx, vz):
        self.resize(len(x))
        for k, a in enumerate((x, z, s, vc, vx, vz)):
            self._buf[k, : self._n] = a


class TwoDFlowField:
    """TwoDFlowField(velX = ConstDef(0.), velZ = ConstDef(0.)), typedefs.jl:20-36"""

    def __init__(self, velX: MotionDef = None, velZ: MotionDef = None):
        self.velX = velX if velX is not None else ConstDef(0.0)
        self.velZ = velZ if velZ is not None else ConstDef(0.0)
        self.u = np.zeros(1)
        self.w = np.zeros(1)
        self.tev = VortList()
        self.lev = VortList()
        self.extv = VortList()


def _init_kinem(kindef: KinemDef, c: float, uref: float) -> KinemPar:
    """Initial KinemPar at t = 0 with the constructor's own typeof chains, typedefs.jl:98-154."""
    k = KinemPar(0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    a, h, u = kindef.alpha, kindef.h, kindef.u
    if type(a) in (EldUpDef, EldRampReturnDef, SinDef, CosDef, FileDef):
        k.alpha = a(0.0)
        k.alphadot = a.deriv(0.0) * uref / c
    elif type(a) is ConstDef:
        k.alpha = a(0.0)
        k.alphadot = 0.0
    if type(h) in (EldUpDef, EldUpIntDef, EldRampReturnDef, SinDef, CosDef):      # typedefs.jl:118-135 (EldUpInttstartDef is not listed there)
        k.h = h(0.0) * c
        k.hdot = h.deriv(0.0) * uref
    elif type(h) is ConstDef:
        k.h = h(0.0) * c
        k.hdot = 0.0
    elif type(h) is FileDef:
        k.h = h(0.0)
        k.hdot = h.deriv(0.0) * uref
    if type(u) is EldUpDef:
        k.u = u(0.0) * uref
        k.udot = u.deriv(0.0) * uref * uref / c
    elif type(u) is ConstDef:
        k.u = u(0.0) * uref
        k.udot = 0.0
    elif type(a) is FileDef:  # sic: the reference tests kindef.alpha here (typedefs.jl:151)
        k.u = u(0.0) * uref
        k.udot = u.deriv(0.0) * uref * uref / c
    return k


class TwoDSurf:
    """TwoDSurf(coord_file, pvt, kindef, lespcrit=zeros(1); c=1., uref=1., ndiv=70, naterm=35,
    initpos=[0.;0.], rho=0.04, camberType="radial"), typedefs.jl:38-176"""

    def __init__(self, coord_file, pvt, kindef, lespcrit=None, *, c=1.0, uref=1.0, ndiv=70, naterm=35,
                 initpos=(0.0, 0.0), rho=0.04, camberType="radial"):
        self.c, self.uref, self.coord_file, self.pvt = float(c), float(uref), coord_file, float(pvt)
        self.ndiv, self.naterm, self.kindef = int(ndiv), int(naterm), kindef
        self.lespcrit = np.zeros(1) if lespcrit is None else np.atleast_1d(np.asarray(lespcrit, dtype=np.float64)).copy()
        self.initpos = np.asarray(initpos, dtype=np.float64)
        self.rho = float(rho)
        theta = np.zeros(ndiv)
        x = np.zeros(ndiv)
        if camberType == "radial":  # :77-82
            dtheta = math.pi / (ndiv - 1)
            for ib in range(ndiv):
                theta[ib] = ib * dtheta
                x[ib] = c / 2.0 * (1 - math.cos(theta[ib]))
        elif camberType == "linear":  # :83-88
            dx = c / (ndiv - 1)
            for ib in range(1, ndiv):
                x[ib] = x[ib - 1] + dx
        cam = np.zeros(ndiv)
        cam_slope = np.zeros(ndiv)
        if coord_file != "FlatPlate":  # :90-92
            from .utils import camber_calc
            cam, cam_slope = camber_calc(x, coord_file)
        self.theta, self.x, self.cam, self.cam_slope = theta, x, np.asarray(cam, float), np.asarray(cam_slope, float)
        self.kinem = _init_kinem(kindef, self.c, self.uref)
        k = self.kinem
        self.bnd_x = np.zeros(ndiv)
        self.bnd_z = np.zeros(ndiv)
        for i in range(ndiv):  # :156-159
            self.bnd_x[i] = -((c - pvt * c) + ((pvt * c - x[i]) * math.cos(k.alpha))) + (self.cam[i] * math.sin(k.alpha)) + self.initpos[0]
            self.bnd_z[i] = k.h + ((pvt * c - x[i]) * math.sin(k.alpha)) + (self.cam[i] * math.cos(k.alpha)) + self.initpos[1]
        self.uind = np.zeros(ndiv)
        self.wind = np.zeros(ndiv)
        self.downwash = np.zeros(ndiv)
        self.a0 = np.zeros(1)
        self.a0dot = np.zeros(1)
        self.aterm = np.zeros(naterm)
        self.adot = np.zeros(naterm)
        self.a0prev = np.zeros(1)
        self.aprev = np.zeros(naterm)
        self.bv = VortList.from_arrays(np.zeros(ndiv - 1), np.zeros(ndiv - 1), np.zeros(ndiv - 1),
                                       np.full(ndiv - 1, 0.02 * c))  # :169-172
        self.levflag = np.zeros(1, dtype=np.int8)


@dataclass
class TwoDOFPar:
    """mutable struct TwoDOFPar, typedefs.jl:178-190"""
    x_alpha: float
    r_alpha: float
    kappa: float
    w_alpha: float
    w_h: float
    w_alphadot: float
    w_hdot: float
    cubic_h_1: float
    cubic_h_3: float
    cubic_alpha_1: float
    cubic_alpha_3: float

    def as_array(self):
        return np.array([self.x_alpha, self.r_alpha, self.kappa, self.w_alpha, self.w_h, self.w_alphadot,
                         self.w_hdot, self.cubic_h_1, self.cubic_h_3, self.cubic_alpha_1, self.cubic_alpha_3])


_K2 = ("alpha", "h", "alphadot", "hdot", "u", "udot", "alphaddot", "hddot", "alpha_pr", "h_pr",
       "alphadot_pr", "alphadot_pr2", "alphadot_pr3", "hdot_pr", "hdot_pr2", "hdot_pr3",
       "alphaddot_pr", "alphaddot_pr2", "alphaddot_pr3", "hddot_pr", "hddot_pr2", "hddot_pr3")


class KinemPar2DOF:
    """mutable struct KinemPar2DOF + constructor, typedefs.jl:193-220"""

    FIELDS = _K2

    def __init__(self, alpha, h, alphadot, hdot, u):
        vals = [alpha, h, alphadot, hdot, u, 0.0, 0.0, 0.0, alpha, h, alphadot, alphadot, alphadot,
                hdot, hdot, hdot, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
        for k, v in zip(_K2, vals):
            setattr(self, k, float(v))

    def as_array(self):
        return np.array([getattr(self, k) for k in _K2], dtype=np.float64)

    def from_array(self, a):
        for k, v in zip(_K2, a):
            setattr(self, k, float(v))
