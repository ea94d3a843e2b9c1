"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import ctypes as C
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c1r_surf, c5_2dof
from synth import s_wake

cabi, L = uf._cabi, uf.lib()
x, z, g, vc = s_wake(3000, uniform_vc=False)
u, w = uf.ind_vel(uf.VortList.from_arrays(x, z, g, vc), x[:100] + 0.01, z[:100])
v = uf.VortList.from_arrays(*s_wake(2500))
uf.mutual_ind(v)
for variant, n in ((1, 1500), (2, 2049), (2, 4100)):
    xx, zz, gg, vv = s_wake(n)
    h = C.c_void_p()
    bx = np.linspace(-1, 0, 69); bz = np.zeros(69); bs = np.full(69, 0.01); bvc = np.full(69, 0.02)
    cabi.check(L.unsflow_wake_create(C.byref(h), n, 69, 0, 1, None, variant))
    cabi.check(L.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (xx, zz, gg, vv, bx, bz, bs, bvc))))
    cabi.check(L.unsflow_wake_step(h, 2, 0.1, 0.0, 0.015))
    cabi.check(L.unsflow_wake_sync(h))
    L.unsflow_wake_destroy(h)
surf, fld, _, dtstar = c1r_surf(uf)
mat, surf, fld = uf.ldvm(surf, fld, 70, dtstar, write_summary=False)
surf, fld, strpar, kin0 = c5_2dof(uf)
mat, surf, fld = uf.ldvm2DOF(surf, fld, strpar, kin0, 50, 0.015, 0, 0, 1000.0, uf.delSpalart(20, 10.0, 1e-6), write_summary=False)
surf, fld, _, dtstar = c1r_surf(uf)
uf.ldvmLin(surf, fld, 40, dtstar, write_summary=False)
surf, fld, _, dtstar = c1r_surf(uf)
uf.lautat(surf, fld, 40, dtstar, write_summary=False)
def mk(x0, phase, lesp):
    kin = uf.KinemDef(uf.SinDef(5 * math.pi / 180, 15 * math.pi / 180, 0.6, phase), uf.ConstDef(0.0), uf.ConstDef(1.0))
    return uf.TwoDSurf("FlatPlate", 0.25, kin, [lesp], initpos=(x0, 0.0))
uf.ldvm([mk(0.0, 0.0, 0.12), mk(1.5, 1.0, 0.2)], uf.TwoDFlowField(), 60, 0.012, write_summary=False)
tabs = np.array([uf.kinematics_table(mk(0.0, 0.3 * k, 0.1), uf.TwoDFlowField(), 40, 0.015) for k in range(5)])
uf.ldvm_batch(mk(0.0, 0.0, 0.1), tabs, np.linspace(0.05, 0.3, 5), 0.015)
print("sanitize_run ok")
