"""A/B of the pair-tile kernel variants on one B200: ms per roll-up step and accuracy against the long-double
truth for each (UNSFLOW_RSQRT_ORDER, UNSFLOW_SYM_VAR) pair.  The settings are read once per process, hence
one subprocess per variant.  Usage: python tools/sym_variants.py [order:var ...]   (default: the whole table)

order 3 = cubic (Halley) rsqrt in FP64 (default), 4 = Halley with the e^2 term in FP32 through F2F
conversions, 5 = same with the double->float step by integer ops, 6 = both conversions by integer ops,
2 = bias-centred Newton (opt-in FAST mode).  var bit 0 = split column-accumulator chains, bit 1 = unroll 2."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NS = [int(v) for v in os.environ.get("SCAN_N", "200000,1000000").split(",")]

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import unsflow_b200 as uf
    from synth import s_wake
    from oracle import oracle as O   # checker
    L, cabi = uf.lib(), uf._cabi
    nbv = 69
    bx = np.linspace(-1.0, 0.0, nbv); bz = np.zeros(nbv); bs = 0.01 * np.ones(nbv); bvc = np.full(nbv, 0.02)
    for n in NS:
        x, z, g, vc = s_wake(n)
        h = C.c_void_p()
        cabi.check(L.unsflow_wake_create(C.byref(h), n, nbv, 0, 1, None, 2))
        cabi.check(L.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc))))
        steps = 3 if n >= 500000 else 10
        cabi.check(L.unsflow_wake_step(h, 2, 0.0, 0.0, 0.0))
        cabi.check(L.unsflow_wake_sync(h))
        best = 1e30
        for rep in range(3):
            cabi.check(L.unsflow_wake_step(h, steps, 0.0, 0.0, 0.0))
            ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
            cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
            best = min(best, msk.value / steps)
        vx, vz = np.zeros(n), np.zeros(n)
        cabi.check(L.unsflow_wake_download(h, None, None, cabi.dptr(vx), cabi.dptr(vz)))
        idx = np.arange(0, n, max(1, n // 24))[:24]
        sx = np.concatenate([x, bx]); sz = np.concatenate([z, bz]); sg = np.concatenate([g, bs]); sv = np.concatenate([vc, bvc])
        tu, tw, au, aw = O.ind_vel_truth(sx, sz, sg, sv, x[idx], z[idx])
        err = max(np.max(np.abs(vx[idx] - tu)) / np.max(au), np.max(np.abs(vz[idx] - tw)) / np.max(aw))
        print(n, best, err, flush=True)
        L.unsflow_wake_destroy(h)
    sys.exit(0)

variants = sys.argv[1:] or ["3:0", "3:1", "3:2", "3:3", "4:0", "4:1", "4:2", "4:3", "5:0", "6:0", "2:0"]
print("| order:var | " + " | ".join(f"N={n} ms (interactions/s) err" for n in NS) + " |")
print("|---|" + "---|" * len(NS))
for v in variants:
    o, var = v.split(":")
    env = dict(os.environ, UNSFLOW_RSQRT_ORDER=o, UNSFLOW_SYM_VAR=var)
    r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print(f"| {v} | failed: {r.stderr.strip()[-200:]} |")
        continue
    cells = []
    for line in r.stdout.strip().splitlines():
        n, ms, err = line.split()
        n, ms, err = int(n), float(ms), float(err)
        cells.append(f"{ms:.3f} ({n * (n + 69.0) / (ms * 1e-3):.4e}) {err:.1e}")
    print(f"| {v} | " + " | ".join(cells) + " |", flush=True)
