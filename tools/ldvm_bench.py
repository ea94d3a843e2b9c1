"""LDVM steps/s vs wake size N on one GPU: config C1 (468 steps from an empty wake) and a
C2-style growth run (pitch-plunge, no deletion) reported per window of steps."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c1_surf, c2_surf
from unsflow_b200.solvers import LdvmHandle, _opts

cabi = uf._cabi
surf, fld, nsteps, dtstar = c1_surf(uf)
for rep in range(2):
    s2, f2, _, _ = c1_surf(uf)
    t0 = time.time()
    mat, s2, f2 = uf.ldvm(s2, f2, nsteps, dtstar, write_summary=False)
    wall = time.time() - t0
    ms, nl = uf.solvers._run.last_timing
    print(f"C1 ldvm {nsteps} steps: device {ms:.2f} ms ({nsteps / ms * 1e3:.0f} steps/s, {ms / nsteps * 1e3:.1f} us/step), "
          f"wall {wall * 1e3:.1f} ms, launches {nl}, Cl_end {mat[-1, 5]:.6f}", flush=True)

total = int(os.environ.get("GROW_STEPS", 20000))
win = int(os.environ.get("GROW_WINDOW", 2000))
surf, fld, dtstar = c2_surf(uf)
h = LdvmHandle(surf, fld, _opts(cabi.DRIVER_LDVM, dtstar, total, uf.delNone(), cabi.SOLVE_EXACT))
h.set_kinematics(uf.kinematics_table(surf, fld, total, dtstar))
done = 0
t0 = time.time()
while done < total:
    h.run(win)
    done += win
    ms, nl = h.timing()
    nt, nlv, _ = h.counts()
    n = nt + nlv
    print(f"C2 growth: steps {done - win}-{done} N={n} (tev {nt} lev {nlv}): {win / ms * 1e3:.0f} steps/s, {ms / win:.3f} ms/step, "
          f"interactions/s {win * float(n) * n / (ms * 1e-3):.3e}", flush=True)
print(f"total wall {time.time() - t0:.1f} s")
h.close()
