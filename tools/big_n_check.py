import sys, time, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import unsflow_b200 as uf
from oracle import oracle as O
from cases import rel_err_normalised
n = 3_000_000
x, z, g, vc = O.s_wake(n)
v = uf.VortList.from_arrays(x, z, g, vc)
t0 = time.time(); uf.mutual_ind(v); t1 = time.time()
idx = np.array([0, 1, 2047, 2048, 1_500_000, 2_999_999, 2_998_000, 777_777])
tu, tw, au, aw = O.ind_vel_truth(x, z, g, vc, x[idx], z[idx])
print("N", n, "wall", t1 - t0, "err", rel_err_normalised(v.vx[idx], tu, au), rel_err_normalised(v.vz[idx], tw, aw),
      "antisym", abs(np.sum(g * v.vx)) / (np.sum(np.abs(g)) * np.max(au)))
