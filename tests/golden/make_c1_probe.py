"""Generates tests/golden/c1_ldvm_probe.json: mat rows of the reference's ldvmLinTest
configuration (test/ldvmLinTest.jl:1-29) run through `ldvm` semantics (solvers.jl:144-268) by
the CPU oracle in exact-solve mode.  The reference itself cannot run here (no julia); the
values agree to all printed digits with the surveyor's independent numpy restatement
(SURVEY.md appendix B)."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import unsflow_b200 as uf
from oracle import oracle as O
from cases import c1_surf

surf, fld, nsteps, dtstar = c1_surf(uf)
tab = uf.kinematics_table(surf, fld, nsteps, dtstar)
sim = O.Sim(surf, fld)
mat = sim.run(O.DRIVER_LDVM, tab, dtstar)
rows = []
for step in (1, 2, 3, 10, 100, 200, 300, 468):
    r = mat[step - 1]
    rows.append(dict(step=step, t=r[0], alpha=r[1], h=r[2], u=r[3], a0=r[4], cl=r[5], cd=r[6], cm=r[7]))
tev = sim.get_vorts(0)
out = dict(config="test/ldvmLinTest.jl:1-29 through ldvm (exact solve)", nsteps=nsteps, dtstar=dtstar, rows=rows,
           ntev=int(len(tev["x"])), sum_gamma_tev=float(tev["s"].sum()),
           oldest_tev=[float(tev["x"][0]), float(tev["z"][0]), float(tev["s"][0])],
           newest_tev=[float(tev["x"][-1]), float(tev["z"][-1]), float(tev["s"][-1])])
json.dump(out, open(os.path.join(HERE, "c1_ldvm_probe.json"), "w"), indent=1)
print("wrote", len(rows), "rows")
