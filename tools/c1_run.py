import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c1_surf
surf, fld, nsteps, dtstar = c1_surf(uf)
n = int(os.environ.get("C1_STEPS", nsteps))
mat, surf, fld = uf.ldvm(surf, fld, n, dtstar, write_summary=False)
print("Cl_end", mat[-1, 5], "timing", uf.solvers._run.last_timing)
