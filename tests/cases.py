"""Shared test configurations."""
import math

import numpy as np


def c1_surf(uf):
    """C1: exactly test/ldvmLinTest.jl:1-29 of the reference (flat plate, 5 deg constant)."""
    alphadef = uf.ConstDef(5.0 * math.pi / 180)
    kin = uf.KinemDef(alphadef, uf.ConstDef(0.0), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [10.15], rho=0.016)
    dtstar = uf.find_tstep(alphadef)
    nsteps = int(round(7.0 / dtstar)) + 1
    return surf, uf.TwoDFlowField(), nsteps, dtstar


def c1r_surf(uf, lespcrit=0.11):
    """C1r: Eldredge ramp to 45 deg, K = 0.2, pivot at LE -> LEV shedding (surveyor's choice)."""
    alphadef = uf.EldUpDef(45.0 * math.pi / 180, 0.2, 0.8)
    kin = uf.KinemDef(alphadef, uf.ConstDef(0.0), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.0, kin, [lespcrit])
    return surf, uf.TwoDFlowField(), 500, uf.find_tstep(alphadef)


def c2_surf(uf, lespcrit=0.18):
    """C2: sinusoidal pitch + cosine plunge (kinem.jl:142-163)."""
    alphadef = uf.SinDef(0.0, 10.0 * math.pi / 180, 0.4, 0.0)
    hdef = uf.CosDef(0.0, 0.1, 0.4, 0.0)
    kin = uf.KinemDef(alphadef, hdef, uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [lespcrit])
    return surf, uf.TwoDFlowField(), uf.find_tstep(alphadef)


def c5_2dof(uf):
    """C5 structural parameters"""
    strpar = uf.TwoDOFPar(0.2, 0.5, 0.05, 1.0, 0.5, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0)
    kin0 = uf.KinemPar2DOF(10.0 * math.pi / 180, 0.0, 0.0, 0.0, 1.0)
    kin = uf.KinemDef(uf.ConstDef(kin0.alpha), uf.ConstDef(0.0), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.35, kin, [0.2])
    return surf, uf.TwoDFlowField(), strpar, kin0


def rel_err_normalised(v, truth, absum):
    """max_i |v_i - truth_i| / max_i sum_j |term_ij|  (SURVEY.md 8c parity metric)"""
    return float(np.max(np.abs(v - truth)) / np.max(absum))
