"""Run under torchrun (one process per GPU): the LDVM stepper with the roll-up pair tiles split
over the ranks must (a) return bit-identical results on every rank and (b) agree with the
single-GPU stepper to round-off.  Also times config-2-like steps at N = 1e5 on G GPUs."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unsflow_b200 as uf
from cases import c2_surf
from synth import s_wake
from unsflow_b200.solvers import LdvmHandle, _opts

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
cabi, L = uf._cabi, uf.lib()
cabi.check(L.unsflow_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = (C.c_ubyte * 128)()
t_uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    cabi.check(L.unsflow_nccl_unique_id(uid))
    t_uid.copy_(torch.tensor(list(uid), dtype=torch.uint8))
dist.broadcast(t_uid, 0)
for i, v in enumerate(t_uid.cpu().tolist()):
    uid[i] = v


def make(nw):
    surf, fld, dtstar = c2_surf(uf)
    x, z, g, vc = s_wake(nw)
    fld.tev = uf.VortList.from_arrays(x[: nw // 2] + 0.5, z[: nw // 2], g[: nw // 2], vc[: nw // 2])
    fld.lev = uf.VortList.from_arrays(x[nw // 2:] + 0.5, z[nw // 2:], g[nw // 2:], vc[nw // 2:])
    return surf, fld, dtstar


nw = int(os.environ.get("MG_WAKE", 100000))
nsteps = int(os.environ.get("MG_STEPS", 20))
surf, fld, dtstar = make(nw)
h = LdvmHandle(surf, fld, _opts(cabi.DRIVER_LDVM, dtstar, nsteps + 3, uf.delNone(), None))
h.set_parallel(rank, world, uid)
peer = h.parallel_info()[2]
h.set_kinematics(uf.kinematics_table(surf, fld, nsteps + 3, dtstar))
m0 = h.run(3)
dist.barrier()
mat = h.run(nsteps)
ms, nl = h.timing()
h.download()
h.close()
digest = hashlib.sha256(np.concatenate([m0.ravel(), mat.ravel(), fld.tev.x, fld.lev.z, surf.aterm]).tobytes()).digest()
t = torch.tensor(list(digest), dtype=torch.uint8, device="cuda")
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    same = all(bool(torch.equal(allt[0], a)) for a in allt)
    # single-GPU stepper on the same input
    s1, f1, _ = make(nw)
    h1 = LdvmHandle(s1, f1, _opts(cabi.DRIVER_LDVM, dtstar, nsteps + 3, uf.delNone(), None))
    h1.set_kinematics(uf.kinematics_table(s1, f1, nsteps + 3, dtstar))
    h1.run(3)
    mat1 = h1.run(nsteps)
    ms1, _ = h1.timing()
    h1.download()
    h1.close()
    dm = float(np.max(np.abs(mat - mat1)))
    dx = float(np.max(np.abs(fld.tev.x - f1.tev.x)))
    print(f"ranks {world} N {nw} (exchange: {'NVLink peer memory, fused finish' if peer else 'NCCL all-reduce'}): bit-identical across ranks {same}; vs single GPU max|dmat| {dm:.2e} max|dx| {dx:.2e}; "
          f"{nsteps} steps: {tm.item():.1f} ms on {world} GPUs ({nsteps / tm.item() * 1e3:.1f} steps/s) vs {ms1:.1f} ms on 1 "
          f"({nsteps / ms1 * 1e3:.1f} steps/s), speed-up {ms1 / tm.item():.2f}")
    assert same and dm <= 1e-9 and dx <= 1e-12 * max(1.0, float(np.max(np.abs(f1.tev.x)))), (same, dm, dx)
dist.barrier()
dist.destroy_process_group()
