"""A few LDVM steps on a pre-seeded wake of N vortices (config C2 regime) -- the command wrapped by the ncu launch
list that names the segments of a stepper step (profiles/r02_stepper_N1e5_launches.txt).
Usage: python tools/stepper_launches.py [N=100000] [steps=6]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c2_surf
from synth import s_wake
from unsflow_b200.solvers import LdvmHandle, _opts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cabi = uf._cabi
surf, fld, dtstar = c2_surf(uf)
x, z, g, vc = s_wake(n)
if os.environ.get("STEPPER_TWO_FAMILIES"):   # half TEVs, half LEVs (the wake of tools/ldvm_multigpu_check.py)
    fld.tev = uf.VortList.from_arrays(x[: n // 2] + 0.5, z[: n // 2], g[: n // 2], vc[: n // 2])
    fld.lev = uf.VortList.from_arrays(x[n // 2:] + 0.5, z[n // 2:], g[n // 2:], vc[n // 2:])
else:
    fld.tev = uf.VortList.from_arrays(x + 0.5, z, g, vc)
h = LdvmHandle(surf, fld, _opts(cabi.DRIVER_LDVM, dtstar, steps, uf.delNone(), None))
h.set_kinematics(uf.kinematics_table(surf, fld, steps, dtstar))
h.run(steps)
ms, nl = h.timing()
print(f"N {n}: {steps} steps in {ms:.3f} ms ({ms / steps * 1e3:.1f} us/step), {nl / steps:.1f} launches/step")
h.close()
