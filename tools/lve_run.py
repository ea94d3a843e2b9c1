import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsflow_b200 as uf
kin = uf.KinemDef(uf.EldUpDef(25 * math.pi / 180, 0.2, 0.8), uf.ConstDef(0.0), uf.ConstDef(1.0))
surf = uf.TwoDSurf("FlatPlate", 0.25, kin, [0.2], rho=1.0)
out = uf.LVE(surf, uf.TwoDFlowField(), int(os.environ.get("LVE_STEPS", 300)), 0.015)
print("LVE", uf.LVE.last_timing, len(out[1].lev))
