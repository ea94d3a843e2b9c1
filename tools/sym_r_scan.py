"""ms per roll-up step of the pair-tile kernel for each tile width (R rows per thread, tile = 256 R) vs N,
next to the direct kernel -- the data behind sym_rows_per_thread() / the pair-tile threshold.
The forced tile width is read once per process (env UNSFLOW_SYM_R), hence one subprocess per R."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NS = [int(v) for v in os.environ.get("SCAN_N", "6000,8000,12000,16000,24000,32000,40000,50000,64000,80000,100000,128000,150000,200000,300000").split(",")]

if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np
    import unsflow_b200 as uf
    from synth import s_wake
    L, cabi = uf.lib(), uf._cabi
    variant = int(sys.argv[1])
    nbv = 69
    bx = np.linspace(-1.0, 0.0, nbv); bz = np.zeros(nbv); bs = 0.01 * np.ones(nbv); bvc = np.full(nbv, 0.02)
    for n in NS:
        x, z, g, vc = s_wake(n)
        h = C.c_void_p()
        cabi.check(L.unsflow_wake_create(C.byref(h), n, nbv, 0, 1, None, variant))
        cabi.check(L.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc))))
        steps = 5 if n >= 100000 else 20
        cabi.check(L.unsflow_wake_step(h, 3, 0.0, 0.0, 0.0))
        cabi.check(L.unsflow_wake_sync(h))
        cabi.check(L.unsflow_wake_step(h, steps, 0.0, 0.0, 0.0))
        ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
        print(n, ms.value / steps, flush=True)
        L.unsflow_wake_destroy(h)
    sys.exit(0)

res = {}
for label, variant, r in (("direct", 1, 0), ("R=2", 2, 2), ("R=4", 2, 4), ("R=8", 2, 8), ("auto", 0, 0)):
    env = dict(os.environ)
    if r:
        env["UNSFLOW_SYM_R"] = str(r)
    out = subprocess.run([sys.executable, __file__, str(variant)], env=env, capture_output=True, text=True, check=True).stdout
    res[label] = {int(l.split()[0]): float(l.split()[1]) for l in out.strip().splitlines()}
cols = list(res)
print("| N | " + " | ".join(f"{c} ms" for c in cols) + " | best (auto / best) |")
print("|---|" + "---|" * (len(cols) + 1))
for n in NS:
    row = [res[c][n] for c in cols]
    best = min(row[:-1])
    print(f"| {n} | " + " | ".join(f"{v:.4f}" for v in row) + f" | {cols[row.index(best)]} ({row[-1] / best:.3f}) |")
