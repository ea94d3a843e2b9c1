"""Host-side mirror of the reference's motion definitions (src/kinem.jl).

These are evaluated on the HOST (in the real drop-in by the reference's own Julia callables)
to tabulate the prescribed motion that the device loop consumes; they are not on the GPU path.
`deriv(t)` is the analytic d/dt the reference obtains with ForwardDiff.derivative
(src/lowOrder2D/calcs.jl:102 ff.).
"""
import math
from dataclasses import dataclass

import numpy as np


class MotionDef:
    """abstract type MotionDef, kinem.jl:1"""

    def __call__(self, t):
        raise NotImplementedError

    def deriv(self, t):
        raise NotImplementedError


@dataclass
class KinemPar:
    """mutable struct KinemPar, kinem.jl:3-10"""
    alpha: float = 0.0
    h: float = 0.0
    alphadot: float = 0.0
    hdot: float = 0.0
    u: float = 0.0
    udot: float = 0.0


@dataclass
class KinemDef:
    """mutable struct KinemDef, kinem.jl:12-16"""
    alpha: MotionDef
    h: MotionDef
    u: MotionDef


@dataclass
class EldUpDef(MotionDef):
    """Eldredge pitch-up ramp, kinem.jl:18-29"""
    amp: float
    K: float
    a: float
    _t1 = 1.0

    def _params(self):
        sm = math.pi * math.pi * self.K / (2 * self.amp * (1 - self.a))
        t1 = self._tstart()
        t2 = t1 + (self.amp / (2 * self.K))
        return sm, t1, t2

    def _tstart(self):
        return 1.0

    def __call__(self, t):
        sm, t1, t2 = self._params()
        return ((self.K / sm) * math.log(math.cosh(sm * (t - t1)) / math.cosh(sm * (t - t2)))) + (self.amp / 2)

    def deriv(self, t):
        sm, t1, t2 = self._params()
        return self.K * (math.tanh(sm * (t - t1)) - math.tanh(sm * (t - t2)))


@dataclass
class EldUptstartDef(EldUpDef):
    """kinem.jl:32-44"""
    tstart: float = 1.0

    def _tstart(self):
        return self.tstart


@dataclass
class EldRampReturnDef(MotionDef):
    """kinem.jl:47-75 (value normalised by the maximum of g over a 0.015-spaced grid)"""
    amp: float
    K: float
    a: float

    def _tstart(self):
        return 1.0

    def _times(self):
        fr = self.K / (math.pi * abs(self.amp))
        t1 = self._tstart()
        t2 = t1 + (1.0 / (2 * math.pi * fr))
        t3 = t2 + ((1 / (4 * fr)) - (1 / (2 * math.pi * fr)))
        t4 = t3 + (1.0 / (2 * math.pi * fr))
        t5 = t4 + t1
        return t1, t2, t3, t4, t5

    def _g(self, t):
        t1, t2, t3, t4, _ = self._times()
        a = self.a
        return np.log((np.cosh(a * (t - t1)) * np.cosh(a * (t - t4))) / (np.cosh(a * (t - t2)) * np.cosh(a * (t - t3))))

    def _maxg(self):
        t5 = self._times()[4]
        nstep = int(round(t5 / 0.015)) + 1
        tt = np.arange(nstep) * 0.015
        return float(np.max(self._g(tt)))

    def __call__(self, t):
        return self.amp * float(self._g(t)) / self._maxg()

    def deriv(self, t):
        t1, t2, t3, t4, _ = self._times()
        a = self.a
        dg = a * (math.tanh(a * (t - t1)) + math.tanh(a * (t - t4)) - math.tanh(a * (t - t2)) - math.tanh(a * (t - t3)))
        return self.amp * dg / self._maxg()


@dataclass
class EldRampReturntstartDef(EldRampReturnDef):
    """kinem.jl:77-106"""
    tstart: float = 1.0

    def _tstart(self):
        return self.tstart


@dataclass
class ConstDef(MotionDef):
    """kinem.jl:108-114"""
    amp: float

    def __call__(self, t):
        return self.amp

    def deriv(self, t):
        return 0.0


@dataclass
class LinearDef(MotionDef):
    """kinem.jl:117-132"""
    tstart: float
    vstart: float
    vend: float
    len: float

    def __call__(self, t):
        if t < self.tstart:
            return self.vstart
        if t > self.tstart + self.len:
            return self.vend
        return self.vstart + (self.vend - self.vstart) / self.len * (t - self.tstart)

    def deriv(self, t):
        if t < self.tstart or t > self.tstart + self.len:
            return 0.0
        return (self.vend - self.vstart) / self.len


@dataclass
class SinDef(MotionDef):
    """kinem.jl:142-151"""
    mean: float
    amp: float
    k: float
    phi: float

    def __call__(self, t):
        return self.mean + self.amp * math.sin(2 * self.k * t + self.phi)

    def deriv(self, t):
        return self.amp * math.cos(2 * self.k * t + self.phi) * 2 * self.k


@dataclass
class CosDef(MotionDef):
    """kinem.jl:154-163"""
    mean: float
    amp: float
    k: float
    phi: float

    def __call__(self, t):
        return self.mean + self.amp * math.cos(2 * self.k * t + self.phi)

    def deriv(self, t):
        return -self.amp * math.sin(2 * self.k * t + self.phi) * 2 * self.k


@dataclass
class StepGustDef(MotionDef):
    """kinem.jl:165-177"""
    amp: float
    tstart: float
    tgust: float

    def __call__(self, t):
        return self.amp if (self.tstart <= t <= self.tstart + self.tgust) else 0.0

    def deriv(self, t):
        return 0.0


@dataclass
class EldUpIntDef(MotionDef):
    """Eldredge ramp integrated in time with a running rectangle rule (plunge from a prescribed plunge-rate ramp),
    kinem.jl:179-213.  `deriv` is what ForwardDiff.derivative returns for that callable: the step count
    round(Int, t/0.015 + 1) carries no derivative, dt = t/(nsteps - 1) does (d dt/dt = 1/(nsteps - 1)), and so do the
    sample times (i-1) dt; for nsteps == 1 the result is the literal 0. (value 0, derivative 0)."""
    amp: float
    K: float
    a: float

    def _tstart(self):
        return 1.0

    def _eval(self, t):
        dt = 0.015
        nsteps = int(round(t / dt + 1))            # round(Int, x): ties to even, like Python's round
        if nsteps == 1:
            return 0.0, 0.0
        dt = t / (nsteps - 1)
        ddt = 1.0 / (nsteps - 1)
        if self.amp == 0.0:
            return 0.0, 0.0
        sm = math.pi * math.pi * self.K / (2 * self.amp * (1 - self.a))
        t1 = self._tstart()
        t2 = t1 + (self.amp / (2 * self.K))
        amp = damp = 0.0
        for i in range(1, nsteps + 1):
            tmpt, dtmpt = (i - 1) * dt, (i - 1) * ddt
            hdot = ((self.K / sm) * math.log(math.cosh(sm * (tmpt - t1)) / math.cosh(sm * (tmpt - t2)))) + (self.amp / 2.0)
            dhdot = self.K * (math.tanh(sm * (tmpt - t1)) - math.tanh(sm * (tmpt - t2))) * dtmpt
            amp = amp + hdot * dt
            damp = damp + dhdot * dt + hdot * ddt
        return amp, damp

    def __call__(self, t):
        return self._eval(t)[0]

    def deriv(self, t):
        return self._eval(t)[1]


@dataclass
class EldUpInttstartDef(EldUpIntDef):
    """kinem.jl:215-249"""
    tstart: float = 1.0

    def _tstart(self):
        return self.tstart


@dataclass
class FileDef(MotionDef):
    """kinem.jl:252-262: nearest-sample lookup (piecewise constant => derivative 0)"""
    tvec: np.ndarray
    var: np.ndarray

    def __call__(self, t):
        return float(self.var[int(np.argmin(np.abs(np.asarray(self.tvec) - t)))])

    def deriv(self, t):
        return 0.0


def eval_kinem(kindef: KinemDef, t: float, c: float, uref: float, prev: KinemPar) -> KinemPar:
    """update_kinem(surf, t), calcs.jl:97-198: dimensional alpha, h, u and rates at time t.
    Types a branch of the reference does not list keep their previous value (the if-chains
    fall through), ConstDef rates are hard-coded 0 (:114, :151, :179), FileDef h and u are not
    scaled by c / uref (:159, :190)."""
    k = KinemPar(prev.alpha, prev.h, prev.alphadot, prev.hdot, prev.u, prev.udot)
    a, h, u = kindef.alpha, kindef.h, kindef.u
    if isinstance(a, (EldUpDef, EldRampReturnDef, ConstDef, SinDef, CosDef, FileDef)):
        k.alpha = a(t)
        k.alphadot = 0.0 if isinstance(a, ConstDef) else a.deriv(t) * uref / c
    if isinstance(h, (EldUpDef, EldUpIntDef, EldRampReturnDef, ConstDef, SinDef, CosDef, FileDef)):      # calcs.jl:131-163
        k.h = h(t) if isinstance(h, FileDef) else h(t) * c
        k.hdot = 0.0 if isinstance(h, ConstDef) else h.deriv(t) * uref
    if isinstance(u, (EldUpDef, EldRampReturnDef, ConstDef, SinDef, CosDef, LinearDef, FileDef)) and not isinstance(
        u, EldUptstartDef
    ):
        k.u = u(t) if isinstance(u, FileDef) else u(t) * uref
        k.udot = 0.0 if isinstance(u, ConstDef) else u.deriv(t) * uref * uref / c
    return k


def eval_externalvel(velX: MotionDef, velZ: MotionDef, t: float, prev_u: float, prev_w: float):
    """update_externalvel(curfield, t), calcs.jl:75-92"""
    u, w = prev_u, prev_w
    if isinstance(velX, (CosDef, SinDef, ConstDef)):
        u = velX(t)
        w = velZ(t)
    if isinstance(velX, StepGustDef):
        u = velX(t)
    if isinstance(velZ, StepGustDef):
        w = velZ(t)
    return u, w
