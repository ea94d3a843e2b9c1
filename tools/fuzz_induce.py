"""Randomised parity sweep of the stateless C-ABI entries (ind_vel, mutual_ind, wakeroll, update_indbound)
against the oracle's long-double truth: random sizes on both sides of every kernel-selection rule, uniform
and per-vortex core radii, empty / tiny / ragged families, pre-loaded vx, vz.  Exit code 1 on any error above
1e-12 (normalised by max_i sum_j |term_ij|)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c1_surf, rel_err_normalised
from oracle import oracle as O

TOL = 1e-12
NCASE = int(os.environ.get("FUZZ_CASES", 120))
rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", 1)))


def rand_n(hi):
    kind = rng.integers(0, 4)
    if kind == 0:
        return int(rng.integers(0, 70))
    if kind == 1:
        return int(rng.choice([255, 256, 257, 511, 512, 513, 2047, 2048, 2049, 16383, 16384, 16385]) if hi > 17000 else rng.integers(1, hi))
    return int(min(hi, np.exp(rng.uniform(np.log(10), np.log(hi)))))


def wake(n, uniform, spread):
    x = spread * rng.random(n) - 0.5 * spread
    z = 0.3 * spread * (rng.random(n) - 0.5)
    g = 0.05 * (rng.random(n) - 0.5)
    vc = np.full(n, 0.02) if uniform else 0.02 * (0.3 + 1.5 * rng.random(n))
    return x, z, g, vc


def sample(n, k=96):
    return np.arange(n) if n <= k else np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, k - 2)]))


worst, bad = 0.0, 0
for case in range(NCASE):
    what = int(rng.integers(0, 4))
    uniform = bool(rng.random() < 0.6)
    if what == 0:       # ind_vel(src, t_x, t_z), calcs.jl:462-477
        ns, nt = rand_n(60000), rand_n(5000)
        x, z, g, vc = wake(ns, uniform, 10.0)
        tx, tz, _, _ = wake(nt, True, 12.0)
        u, w = uf.ind_vel(uf.VortList.from_arrays(x, z, g, vc), tx, tz)
        idx = sample(nt)
        if nt == 0 or ns == 0:
            err = float(np.max(np.abs(u), initial=0.0) + np.max(np.abs(w), initial=0.0))
        else:
            tu, tw, au, aw = O.ind_vel_truth(x, z, g, vc, tx[idx], tz[idx])
            err = max(rel_err_normalised(u[idx], tu, au), rel_err_normalised(w[idx], tw, aw))
        desc = f"ind_vel ns {ns} nt {nt}"
    elif what == 1:     # mutual_ind(vorts), calcs.jl:491-510 (accumulates)
        n = max(2, rand_n(70000))
        x, z, g, vc = wake(n, uniform, 0.02 * n + 1.0)
        v = uf.VortList.from_arrays(x, z, g, vc)
        v.vx[:] = 0.25; v.vz[:] = -1.5
        uf.mutual_ind(v)
        idx = sample(n)
        tu, tw, au, aw = O.ind_vel_truth(x, z, g, vc, x[idx], z[idx])
        err = max(rel_err_normalised(v.vx[idx] - 0.25, tu, au), rel_err_normalised(v.vz[idx] + 1.5, tw, aw))
        desc = f"mutual_ind n {n}"
    else:               # wakeroll (calcs.jl:222-294) + update_indbound (:35-38) with ragged families
        surf, fld, _, dtstar = c1_surf(uf)
        nt_, nl_, ne_ = rand_n(30000), rand_n(3000), rand_n(300)
        if what == 3:
            nt_, nl_, ne_ = rand_n(3000), rand_n(300), 0
        x, z, g, vc = wake(nt_ + nl_ + ne_, uniform, 8.0)
        x = x - x.max() - 0.2 if len(x) else x
        fld.tev = uf.VortList.from_arrays(x[:nt_], z[:nt_], g[:nt_], vc[:nt_])
        fld.lev = uf.VortList.from_arrays(x[nt_:nt_ + nl_], z[nt_:nt_ + nl_], g[nt_:nt_ + nl_], vc[nt_:nt_ + nl_])
        fld.extv = uf.VortList.from_arrays(x[nt_ + nl_:], z[nt_ + nl_:], g[nt_ + nl_:], vc[nt_ + nl_:])
        fld.u[0], fld.w[0] = 0.07, -0.03
        surf.bv.s[:] = 0.01 * rng.standard_normal(surf.ndiv - 1)
        surf.bv.x[:] = 0.5 * (surf.bnd_x[1:] + surf.bnd_x[:-1]); surf.bv.z[:] = 0.5 * (surf.bnd_z[1:] + surf.bnd_z[:-1])
        sx, sz, sg, svc = (np.concatenate([a, b]) for a, b in ((x, surf.bv.x), (z, surf.bv.z), (g, surf.bv.s), (vc, surf.bv.vc)))
        uf.update_indbound(surf, fld)
        err = 0.0
        if len(x):
            tu, tw, au, aw = O.ind_vel_truth(x, z, g, vc, surf.bnd_x, surf.bnd_z)
            err = max(rel_err_normalised(surf.uind, tu, au), rel_err_normalised(surf.wind, tw, aw))
        x0, z0 = x.copy(), z.copy()
        uf.wakeroll(surf, fld, 0.015)
        if len(x):
            idx = sample(len(x))
            tu, tw, au, aw = O.ind_vel_truth(sx, sz, sg, svc, x0[idx], z0[idx])
            vx = np.concatenate([fld.tev.vx, fld.lev.vx, fld.extv.vx])[idx]
            vz = np.concatenate([fld.tev.vz, fld.lev.vz, fld.extv.vz])[idx]
            xn = np.concatenate([fld.tev.x, fld.lev.x, fld.extv.x])[idx]
            err = max(err, rel_err_normalised(vx - 0.07, tu, au), rel_err_normalised(vz + 0.03, tw, aw),
                      float(np.max(np.abs(xn - (x0[idx] + 0.015 * vx)))) )
        desc = f"wakeroll+indbound tev {nt_} lev {nl_} extv {ne_}"
    worst = max(worst, err)
    bad += err > TOL
    print(f"case {case:3d} {desc} uniform {uniform} err {err:.2e}{'' if err <= TOL else '  <-- FAIL'}", flush=True)
print(f"{NCASE} cases, worst normalised error {worst:.2e}, {bad} above {TOL}")
sys.exit(1 if bad else 0)
