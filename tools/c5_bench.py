"""BASELINE config 5: ldvm2DOF aeroelastic pitch-plunge coupling with delSpalart merging on a
1e5-vortex wake (wake pre-seeded with S-wake(1e5) as TEVs so merging is active from step 1)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c5_2dof
from synth import s_wake
from unsflow_b200.solvers import LdvmHandle, _opts

cabi = uf._cabi
nw = int(os.environ.get("C5_WAKE", 100000))
nsteps = int(os.environ.get("C5_STEPS", 200))
surf, fld, strpar, kin0 = c5_2dof(uf)
x, z, g, vc = s_wake(nw)
fld.tev = uf.VortList.from_arrays(x + 0.5, z, g, vc)
dv = uf.delSpalart(nw, 10.0, 1e-6)          # limit = initial wake size: merging from step 1
opts = _opts(cabi.DRIVER_LDVM2DOF, 0.015, nsteps + 5, dv, cabi.SOLVE_EXACT)
h = LdvmHandle(surf, fld, opts)
h.set_2dof(strpar, kin0)
h.set_kinematics(uf.kinematics_table(surf, fld, nsteps + 5, 0.015, two_dof=True))
h.run(5)
t0 = time.time()
mat = h.run(nsteps)
wall = time.time() - t0
ms, nl = h.timing()
nt, nlv, _ = h.counts()
h.get_2dof(kin0)
print(f"C5 ldvm2DOF + delSpalart: wake {nw} -> tev {nt} lev {nlv} after {nsteps + 5} steps; {nsteps} timed steps: "
      f"{ms:.1f} ms device ({nsteps / ms * 1e3:.1f} steps/s, {ms / nsteps:.3f} ms/step), wall {wall:.2f} s, launches/step {nl / nsteps:.1f}")
print(f"alpha_end {mat[-1, 1]:.6f} h_end {mat[-1, 2]:.6f} Cl_end {mat[-1, 5]:.6f} finite {bool(np.all(np.isfinite(mat)))} "
      f"merged {nw + nsteps + 5 - nt} TEVs")
h.close()
