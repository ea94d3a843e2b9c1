"""Bound-point sweep kernel (k_sweep) alone: time vs wake size, 70 targets, device-resident sources.
UNSFLOW_SWEEP_M forces the CTAs per SM.  Usage: python tools/sweep_probe.py [ndiv=70]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsflow_b200 as uf

L = uf.lib()
cabi = uf._cabi
ndiv = int(sys.argv[1]) if len(sys.argv) > 1 else 70
tf, mhz = C.c_double(), C.c_double()
cabi.check(L.unsflow_fp64_peak(C.byref(tf), C.byref(mhz)))
print(f"UNSFLOW_SWEEP_M={os.environ.get('UNSFLOW_SWEEP_M', '(rule)')}  ndiv={ndiv}  measured FP64 peak {tf.value:.2f} TF")
print("| N sources | k_sweep us | sources/s | alg. TFLOP/s | of measured FP64 peak | generic us |")
print("|---|---|---|---|---|---|")
for n in (500, 2000, 10000, 30000, 100000, 300000, 1000000, 4000000):
    a, b = C.c_float(), C.c_float()
    cabi.check(L.unsflow_bench_sweep(n, ndiv, 10, C.byref(a), C.byref(b)))
    rate = n / (a.value * 1e-3)
    print(f"| {n} | {a.value * 1e3:.2f} | {rate:.3e} | {13 * ndiv * rate / 1e12:.2f} | {13 * ndiv * rate / 1e12 / tf.value:.3f} | {b.value * 1e3:.1f} |", flush=True)
