"""Run under torchrun (one process per GPU): the multi-GPU split BASELINE.json names -- the free vortices are cut
into contiguous TARGET slices, every rank evaluates its slice against all sources with the one-sided kernel (the only
multi-GPU path for non-uniform core radii) and the convected positions are exchanged with an NCCL all-gather each
step (wake.cu, kernel variant 1).  Checks, on a non-uniform-core field:
  (a) every rank's velocities of its own slice against the oracle's long-double truth (<= 1e-12 normalised),
  (b) after the exchange all ranks hold bit-identical positions,
  (c) positions after `steps` steps equal those of a 1-rank run of the same kernel to round-off,
and times the steps (max over ranks)."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsflow_b200 as uf
from oracle import oracle as O      # checker
from synth import s_wake

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

n = int(os.environ.get("AG_N", 60000))
steps = int(os.environ.get("AG_STEPS", 3))
uniform = os.environ.get("AG_UNIFORM", "0") == "1"
nbv, dt = 69, 0.015
x, z, g, vc = s_wake(n, uniform_vc=uniform)
bx = np.linspace(-1.0, 0.0, nbv); bz = 0.05 * np.ones(nbv); bs = 0.01 * np.cos(np.arange(nbv)); bvc = np.full(nbv, 0.02)
up = [cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc)]


def run(nranks, r):
    h = C.c_void_p()
    cabi.check(L.unsflow_wake_create(C.byref(h), n, nbv, r, nranks, uid if nranks > 1 else None, 1))
    cabi.check(L.unsflow_wake_upload(h, *up))
    cabi.check(L.unsflow_wake_step(h, 1, 0.1, 0.0, 0.0))          # dt = 0: velocities only
    vx, vz, gx, gz = (np.zeros(n) for _ in range(4))
    cabi.check(L.unsflow_wake_download(h, None, None, cabi.dptr(vx), cabi.dptr(vz)))
    cabi.check(L.unsflow_wake_step(h, steps, 0.1, 0.0, dt))
    ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
    cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))
    cabi.check(L.unsflow_wake_download(h, cabi.dptr(gx), cabi.dptr(gz), None, None))
    L.unsflow_wake_destroy(h)
    return vx, vz, gx, gz, ms.value


vx, vz, gx, gz, ms = run(world, rank)
b, e = C.c_int64(), C.c_int64()
cabi.check(L.unsflow_partition(n, world, rank, 1, C.byref(b), C.byref(e)))
idx = np.linspace(b.value, e.value - 1, 16).astype(np.int64)
sx = np.concatenate([x, bx]); sz = np.concatenate([z, bz]); sg = np.concatenate([g, bs]); sv = np.concatenate([vc, bvc])
tu, tw, au, aw = O.ind_vel_truth(sx, sz, sg, sv, x[idx], z[idx])
err = max(np.max(np.abs(vx[idx] - 0.1 - tu)) / np.max(au), np.max(np.abs(vz[idx] - tw)) / np.max(aw))      # 0.1 = free stream
dig = hashlib.sha256(gx.tobytes() + gz.tobytes()).digest()
td = torch.tensor(list(dig), dtype=torch.uint8, device="cuda")
allg = [torch.zeros_like(td) for _ in range(world)]
dist.all_gather(allg, td)
tt = torch.tensor([err, ms], dtype=torch.float64, device="cuda")
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    same = all(bool(torch.equal(allg[0], a)) for a in allg)
    _, _, gx1, gz1, ms1 = run(1, 0)
    dx = float(max(np.max(np.abs(gx - gx1)), np.max(np.abs(gz - gz1))))
    inter = float(n) * (n + nbv) * steps
    print(f"all-gather split: ranks {world} N {n} ({'uniform' if uniform else 'per-vortex'} core radii): slice velocities vs long-double truth "
          f"max normalised err over ranks {tt[0].item():.2e}; positions bit-identical across ranks {same}; vs 1 rank after {steps} steps "
          f"max|dx| {dx:.2e}; {tt[1].item() / steps:.3f} ms/step on {world} GPUs ({inter / (tt[1].item() * 1e-3):.3e} interactions/s) vs "
          f"{ms1 / steps:.3f} ms/step on 1 (speed-up {ms1 / tt[1].item():.2f})")
    assert same and tt[0].item() <= 1e-12 and dx <= 1e-12
dist.barrier()
dist.destroy_process_group()
