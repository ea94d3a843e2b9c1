"""max |GPU - oracle| of the Cl/Cd/Cm histories in both shedding-solve modes on the configurations of the parity tests
(debug aid: which tolerance does NLsolve mode need, and where does the difference start?)"""
import copy
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c1_surf, c1r_surf, c2_surf
from oracle import oracle as O


def attached():
    kin = uf.KinemDef(uf.SinDef(2 * math.pi / 180, 4 * math.pi / 180, 0.3, 0.0), uf.CosDef(0.0, 0.05, 0.3, 0.0), uf.ConstDef(1.0))
    return uf.TwoDSurf("FlatPlate", 0.3, kin, [5.0]), uf.TwoDFlowField(), 500, 0.015


cases = {"c1": lambda: c1_surf(uf), "c1r150": lambda: c1r_surf(uf)[:2] + (150, c1r_surf(uf)[3]), "c2_100": lambda: c2_surf(uf)[:2] + (100, c2_surf(uf)[2]),
         "attached500": attached}
for name, mk in cases.items():
    for mode in (0, 1):
        surf, fld, n, dtstar = mk()
        dt = dtstar * surf.c / surf.uref
        tab = uf.kinematics_table(surf, fld, n, dt)
        sim = O.Sim(copy.deepcopy(surf), copy.deepcopy(fld))
        om = sim.run(O.DRIVER_LDVM, tab, dt, solver_mode=mode)
        gm, s2, f2 = uf.ldvm(surf, fld, n, dtstar, write_summary=False, solve_mode=mode)
        d = np.abs(gm - om)
        per_step = d[:, 4:8].max(axis=1)
        first = int(np.argmax(per_step > 1e-10)) if np.any(per_step > 1e-10) else -1
        print(f"{name:12s} mode {mode}: max|dmat| {d.max():.2e} (a0 {d[:,4].max():.2e} cl {d[:,5].max():.2e} cd {d[:,6].max():.2e} cm {d[:,7].max():.2e}) "
              f"first step above 1e-10: {first}; at step 10: {per_step[min(10, n - 1)]:.2e}; nlev {len(f2.lev)}", flush=True)
