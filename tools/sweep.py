"""interactions/s vs wake N (BASELINE metric) for the direct and pair-tile kernels, one GPU."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsflow_b200 as uf
from synth import s_wake

L = uf.lib()
cabi = uf._cabi
tf, mhz = C.c_double(), C.c_double()
cabi.check(L.unsflow_fp64_peak(C.byref(tf), C.byref(mhz)))
print(f"| N | kernel | ms/step | interactions/s | alg. TFLOP/s (13 flop) | of measured FP64 peak {tf.value:.2f} TF |")
print("|---|---|---|---|---|---|")
nbv = 69
bx = np.linspace(-1.0, 0.0, nbv); bz = np.zeros(nbv); bs = 0.01 * np.ones(nbv); bvc = np.full(nbv, 0.02)
for n in (1000, 3000, 10000, 30000, 100000, 300000, 1000000):
    x, z, g, vc = s_wake(n)
    for variant, name in ((1, "direct"), (2, "pair-tile")):
        if variant == 2 and n < 10000:
            continue
        h = C.c_void_p()
        cabi.check(L.unsflow_wake_create(C.byref(h), n, nbv, 0, 1, None, variant))
        cabi.check(L.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc))))
        steps = 3 if n >= 300000 else 20
        cabi.check(L.unsflow_wake_step(h, 3, 0.0, 0.0, 0.0))
        cabi.check(L.unsflow_wake_sync(h))
        cabi.check(L.unsflow_wake_step(h, steps, 0.0, 0.0, 0.0))
        ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
        rate = float(n) * (n + nbv) * steps / (ms.value * 1e-3)
        print(f"| {n} | {name} | {ms.value / steps:.3f} | {rate:.3e} | {13 * rate / 1e12:.2f} | {13 * rate / 1e12 / tf.value:.3f} |", flush=True)
        L.unsflow_wake_destroy(h)

# ---- bound-point sweep (update_indbound): 70 targets x N sources, device-resident ---------------
print()
print("| N sources | kernel | ms | sources/s | HBM GB/s (24 B/source) | alg. TFLOP/s | of measured FP64 peak |")
print("|---|---|---|---|---|---|---|")
for n in (10000, 100000, 1000000, 4000000):
    a, b = C.c_float(), C.c_float()
    cabi.check(L.unsflow_bench_sweep(n, 70, 5, C.byref(a), C.byref(b)))
    for name, ms in (("k_sweep", a.value), ("generic direct", b.value)):
        rate = n / (ms * 1e-3)
        print(f"| {n} | {name} | {ms:.4f} | {rate:.3e} | {24 * rate / 1e9:.1f} | {13 * 70 * rate / 1e12:.2f} | {13 * 70 * rate / 1e12 / tf.value:.3f} |", flush=True)
