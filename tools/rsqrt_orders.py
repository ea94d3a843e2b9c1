import ctypes as C, sys
sys.path.insert(0,'/root/repo')
import unsflow_b200 as uf
L, cabi = uf.lib(), uf._cabi
for o in (2,3,4,5,6,7):
    e = C.c_double()
    cabi.check(L.unsflow_rsqrt_probe_order(o, 1<<26, 1e-8, 1e18, C.byref(e)))
    print("order", o, "max rel err", e.value)
