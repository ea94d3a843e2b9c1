"""Host-side utilities mirrored from src/utils.jl (not on the GPU path)."""
import numpy as np

from .kinem import (ConstDef, CosDef, EldRampReturnDef, EldUpDef, SinDef)


def find_tstep(kin):
    """find_tstep, utils.jl:52-102"""
    if isinstance(kin, (CosDef, SinDef)):
        return min(0.015 * 0.2 / (kin.k * kin.amp), 15.0)
    if isinstance(kin, (EldUpDef, EldRampReturnDef)):
        return min(0.015 * 0.2 / kin.K, 0.015)
    if isinstance(kin, ConstDef):
        return 0.015
    raise TypeError(f"find_tstep: no method for {type(kin).__name__}")


def simpleTrapz(y, x):
    """simpleTrapz, utils.jl:105-115 (left-to-right accumulation)"""
    if len(y) != len(x):
        raise ValueError("Vectors must be of same length")
    r = 0.0
    for i in range(1, len(y)):
        r += (x[i] - x[i - 1]) * (y[i] + y[i - 1])
    return r / 2.0


def camber_calc(x, airfoil):
    """camber_calc, utils.jl:118-150.  Geometry pre-processing (Dierckx splines in the
    reference, FITPACK through scipy here -- the same library); host only."""
    from scipy.interpolate import UnivariateSpline

    x = np.asarray(x, dtype=np.float64)
    ndiv = len(x)
    c = x[-1]
    in_air = np.loadtxt(airfoil)
    xcoord, ycoord = in_air[:, 0], in_air[:, 1]
    xsum = np.concatenate([[0.0], np.cumsum(np.abs(np.diff(xcoord)))])
    y_spl = UnivariateSpline(xsum, ycoord, k=3, s=0)
    y_ans = np.zeros(2 * ndiv)
    y_ans[:ndiv] = y_spl(x / c)
    y_ans[ndiv:] = y_spl(x[-1] / c + x / c)
    cam = np.array([(y_ans[i] + y_ans[2 * ndiv - 1 - i]) * c / 2 for i in range(ndiv - 1, -1, -1)])
    cam[0] = 0.0
    cam_spl = UnivariateSpline(x, cam, k=3, s=0)
    return cam, cam_spl.derivative()(x)
