"""BASELINE config 3: 4096 independent LDVM kinematics/LESPcrit sweeps (64 reduced frequencies x
64 LESP thresholds, SinDef pitch amplitude 20 deg, 500 steps each), one CTA per member."""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsflow_b200 as uf

if "LOCAL_RANK" in os.environ:
    uf._cabi.check(uf._cabi.lib().unsflow_set_device(int(os.environ["LOCAL_RANK"])))
nk = int(os.environ.get("ENS_NK", 64))
nl = int(os.environ.get("ENS_NL", 64))
nsteps = int(os.environ.get("ENS_STEPS", 500))
dtstar = 0.015
tables, lesp = [], []
base = None
t0 = time.time()
for k in np.linspace(0.05, 1.0, nk):
    kin = uf.KinemDef(uf.SinDef(0.0, 20.0 * math.pi / 180, float(k), 0.0), uf.ConstDef(0.0), uf.ConstDef(1.0))
    surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [0.1])
    base = base or surf
    tab = uf.kinematics_table(surf, uf.TwoDFlowField(), nsteps, dtstar)
    for l in np.linspace(0.05, 0.35, nl):
        tables.append(tab)
        lesp.append(float(l))
world = int(os.environ.get("ENS_WORLD", 1))      # emulate rank 0's shard of a `world`-GPU run on one GPU
tables = np.array(tables)
if world > 1:
    mine = uf.ensemble_shard(tables, lesp, int(os.environ.get("ENS_RANK", os.environ.get("LOCAL_RANK", 0))), world) if os.environ.get("ENS_DEAL", "1") == "1" else np.arange(len(lesp))[0::world]
    tables, lesp = tables[mine], list(np.array(lesp)[mine])
print(f"members {len(lesp)} steps {nsteps} (tables built in {time.time() - t0:.1f} s)", flush=True)
t0 = time.time()
mats, ntev, nlev, ms = uf.ldvm_batch(base, tables, np.array(lesp), dtstar)
wall = time.time() - t0
nvort = ntev + nlev
# interactions actually evaluated: per step N(N+69) roll-up + 70 N sweep, N grows each step
print(f"kernel {ms:.1f} ms, wall {wall:.2f} s, member-steps/s {len(lesp) * nsteps / (ms * 1e-3):.3e}, "
      f"final N min/mean/max {nvort.min()}/{nvort.mean():.0f}/{nvort.max()}, finite {bool(np.all(np.isfinite(mats)))}")
print("Cl range at last step", float(mats[:, -1, 5].min()), float(mats[:, -1, 5].max()))
