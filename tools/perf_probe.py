"""Quick on-GPU probe (not the bench): FP64 peak, rsqrt accuracy, wake-step rate vs N.
(Accuracy against the oracle is checked by tests/ and by bench.py's verify step, not here.)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsflow_b200 as uf
from synth import s_wake

L = uf.lib()
cabi = uf._cabi
tf, mhz = C.c_double(), C.c_double()
cabi.check(L.unsflow_fp64_peak(C.byref(tf), C.byref(mhz)))
print(f"fp64_peak_tflops {tf.value:.3f} est_mhz {mhz.value:.0f}", flush=True)
e0, e2, e3 = C.c_double(), C.c_double(), C.c_double()
cabi.check(L.unsflow_rsqrt_probe(1 << 26, 1e-8, 1e18, C.byref(e0), C.byref(e2), C.byref(e3)))
print(f"rsqrt_rel_err seed {e0.value:.4e} newton {e2.value:.4e} cubic {e3.value:.4e}", flush=True)

variants = [int(v) for v in os.environ.get("PROBE_VARIANTS", "1").split(",")]
for n in [int(s) for s in os.environ.get("PROBE_N", "10000,100000,1000000").split(",")]:
    x, z, g, vc = s_wake(n)
    nbv = 69
    bx = np.linspace(-1.0, 0.0, nbv); bz = np.zeros(nbv); bs = 0.01 * np.ones(nbv); bvc = np.full(nbv, 0.02)
    for variant in variants:
        h = C.c_void_p()
        cabi.check(L.unsflow_wake_create(C.byref(h), n, nbv, 0, 1, None, variant))
        cabi.check(L.unsflow_wake_upload(h, *(cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc))))
        steps = 3 if n >= 500000 else 10
        cabi.check(L.unsflow_wake_step(h, 2, 0.0, 0.0, 0.0))
        cabi.check(L.unsflow_wake_sync(h))
        cabi.check(L.unsflow_wake_step(h, steps, 0.0, 0.0, 0.0))
        ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
        inter = float(n) * (n + nbv) * steps
        rate = inter / (msk.value * 1e-3)
        print(f"N {n} variant {variant} steps {steps} ms_total {ms.value:.3f} ms_induce {msk.value:.3f} "
              f"interactions/s {rate:.4e} alg_TF {13 * rate / 1e12:.2f} frac_of_measured_peak {13 * rate / 1e12 / tf.value:.3f}",
              flush=True)
        L.unsflow_wake_destroy(h)
