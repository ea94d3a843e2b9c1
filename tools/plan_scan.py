"""ms per roll-up step of the direct kernel for forced tilings (env UNSFLOW_PLAN="R,ts,mult") vs N."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NS = [int(v) for v in os.environ.get("SCAN_N", "1000,2000,4000,6000,8000,12000,16000").split(",")]
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np
    import unsflow_b200 as uf
    from synth import s_wake
    L, cabi = uf.lib(), uf._cabi
    nbv = 69
    bx = np.linspace(-1.0, 0.0, nbv); bz = np.zeros(nbv); bs = 0.01 * np.ones(nbv); bvc = np.full(nbv, 0.02)
    for n in NS:
        x, z, g, vc = s_wake(n)
        h = C.c_void_p()
        cabi.check(L.unsflow_wake_create(C.byref(h), n, nbv, 0, 1, None, 1))
        cabi.check(L.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc))))
        cabi.check(L.unsflow_wake_step(h, 5, 0.0, 0.0, 0.0))
        cabi.check(L.unsflow_wake_sync(h))
        cabi.check(L.unsflow_wake_step(h, 50, 0.0, 0.0, 0.0))
        ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
        print(n, ms.value / 50, flush=True)
        L.unsflow_wake_destroy(h)
    sys.exit(0)
plans = ["auto"] + [f"{r},{ts},{m}" for r, ts in ((1, 64), (1, 512), (2, 512), (4, 512)) for m in (2, 4, 8)]
res = {}
for pl in plans:
    env = dict(os.environ)
    if pl != "auto":
        env["UNSFLOW_PLAN"] = pl
    out = subprocess.run([sys.executable, __file__, "x"], env=env, capture_output=True, text=True)
    if out.returncode:
        print(pl, "failed:", out.stderr[-300:]); continue
    res[pl] = {int(l.split()[0]): float(l.split()[1]) for l in out.stdout.strip().splitlines()}
cols = list(res)
print("| N | " + " | ".join(cols) + " | best |")
print("|---|" + "---|" * (len(cols) + 1))
for n in NS:
    row = [res[c][n] for c in cols]
    print(f"| {n} | " + " | ".join(f"{v * 1e3:.1f}" for v in row) + f" | {cols[row.index(min(row))]} |")
