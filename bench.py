#!/usr/bin/env python
"""bench.py -- headline benchmark of the UNSflow discrete-vortex hot path on B200.

Metric (BASELINE.json): FP64 vortex interactions/s.  One *interaction* = one ordered
(target <- source) evaluation of the Vatistas kernel; 13 algorithmic flop each.

Workload (config.workload): C4 -- one `wakeroll` time step (src/lowOrder2D/calcs.jl:222-294:
all-pairs induction among N = 1e6 free vortices + 69 bound vortices, free stream, explicit
Euler convection) on the synthetic S-wake(1e6) field, kept device-resident.  With --gpus G > 1
(launched by torchrun, one process per GPU) the pair work is split over ranks and the step ends
with the NCCL exchange; total work is fixed => "scaling": "strong".

  value     whole-job interactions/s with state resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the C-ABI with HOST buffers: upload of the vortex state from
            host memory, one step, download of positions+velocities, all inside the timed region
  roofline  the dominant kernel against the FP64 (non-tensor) DFMA peak MEASURED in this run
  cpu_baseline  the reference's own loop (oracle port of mutual_ind) on one host core, bounded sample

`--impl reference` times the reference CPU path (oracle port, all host threads) instead.
"""
import argparse
import ctypes as C
import json
import copy
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 13.0
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_induce_sym<8,3> launch at N = 1e6 from the `ncu --set full`
# capture of this round's final build (profiles/r02_sym_R8_N1e6_ncu_full.txt: 4.488 GB read + 5.494 GB written).  It is
# a constant taken from that capture, not measured in this run (ncu cannot run inside the bench); algorithmic bytes
# are 64 MB -- the rest is the read-modify-write of the per-CTA partial planes after every pair tile (2 x 32 KB x
# 119 805 tiles + the zeroing of the touched blocks), 0.25 % of DRAM bandwidth: the kernel is FP64-bound.
NCU_TRAFFIC_SYM8_N1E6 = 9.982e9
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 148 SM x 64 DFMA lanes x 2 flop x 1.965 GHz = 37.2
NBV = 69                        # ndiv - 1 bound vortices (typedefs.jl:68)


def env_int(k, d):
    return int(os.environ.get(k, d))


# stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner,
# torchrun notices) are redirected to stderr for the life of the process.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_baseline_port(O, n_sample=65536):
    """reference loop (mutual_ind, calcs.jl:491-510) on ONE host core, bounded sample of the
    same S-wake field (the reference is single-threaded)"""
    x, z, g, vc = O.s_wake(1_000_000)
    t, _, _ = O.mutual_ind_time(x[:n_sample], z[:n_sample], g[:n_sample], vc[:n_sample])
    inter = float(n_sample) * (n_sample - 1)
    return {"value": inter / t, "unit": "interactions/s", "cores": 1, "kind": "port",
            "sample": f"oracle port of mutual_ind (symmetric i<j loop, 2 sqrt + 2 div per pair) on the first {n_sample} "
                      f"vortices of S-wake(1e6): {inter:.3e} interactions in {t:.1f} s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Julia
    reference cannot run here) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    n = args.n
    x, z, g, vc = O.s_wake(n)
    nth = O.num_threads()
    nt = 48 * nth                      # targets per step: ~0.3 s of work per thread-step at 1.8e8/s
    sx = np.concatenate([x, np.linspace(-1, 0, NBV)])
    sz = np.concatenate([z, np.zeros(NBV)])
    sg = np.concatenate([g, np.full(NBV, 0.01)])
    sv = np.concatenate([vc, np.full(NBV, 0.02)])
    times = []
    for it in range(args.warmup + args.steps):
        lo = (it * nt) % (n - nt)
        t, _, _ = O.ind_vel_slice_mt(sx, sz, sg, sv, x[lo:lo + nt], z[lo:lo + nt])
        if it >= args.warmup:
            times.append(t)
    inter = float(nt) * (n + NBV)
    T = sum(times)
    val = inter * len(times) / T
    cb = {"value": val, "unit": "interactions/s", "cores": nth, "kind": "port",
          "sample": f"{nt} targets x {n + NBV} sources per step (ind_vel expression of calcs.jl:472-473, target-sliced over "
                    f"{nth} pthreads); the reference itself is single-threaded Julia and cannot run here"}
    line = {"impl": "reference", "metric": "FP64 vortex interactions/s", "value": val, "unit": "interactions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4 wakeroll step, S-wake(N={n}) + {NBV} bound vortices", "n_vortices": n,
                       "note": "bounded target sample per step; CPU oracle port"},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="free vortices (C4: 1e6)")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 direct, 2 symmetric")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ldvm-steps", type=int, default=20, help="extra: LDVM steps/s on a 1e5 wake (0 = skip)")
    ap.add_argument("--ensemble-members", type=int, default=4096, help="extra: config C3 ensemble (0 = skip)")
    ap.add_argument("--ensemble-steps", type=int, default=500)
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import unsflow_b200 as uf
    from synth import s_wake
    from oracle import oracle as O   # checker only: cpu_baseline leg and the verify step

    cabi = uf._cabi
    L = uf.lib()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    cabi.check(L.unsflow_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, K, W = args.n, args.steps, args.warmup
    x, z, g, vc = s_wake(n)
    bx = np.linspace(-1.0, 0.0, NBV); bz = np.zeros(NBV); bs = np.full(NBV, 0.01); bvc = np.full(NBV, 0.02)
    dt, uinf = 0.015, 1.0

    uid = (C.c_ubyte * 128)()
    if world > 1:
        t_uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            cabi.check(L.unsflow_nccl_unique_id(uid))
            t_uid.copy_(torch.tensor(list(uid), dtype=torch.uint8))
        dist.broadcast(t_uid, 0)
        for i, v in enumerate(t_uid.cpu().tolist()):
            uid[i] = v
    h = C.c_void_p()
    cabi.check(L.unsflow_wake_create(C.byref(h), n, NBV, rank, world, uid if world > 1 else None, args.variant))
    up = [cabi.dptr(a) for a in (x, z, g, vc, bx, bz, bs, bvc)]
    cabi.check(L.unsflow_wake_upload(h, *up))

    # measured FP64 peak (MEASURED_PEAKS.json has no FP64 figure): DFMA-chain microbenchmark
    tf, mhz = C.c_double(), C.c_double()
    cabi.check(L.unsflow_fp64_peak(C.byref(tf), C.byref(mhz)))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        flush.zero_()
        torch.cuda.synchronize()
        cabi.check(L.unsflow_wake_step(h, 1, uinf, 0.0, dt))
        ms, msk, nl = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h, C.byref(ms), C.byref(msk), C.byref(nl)))   # synchronises
        return ms.value, msk.value, nl.value

    for _ in range(W):
        one_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    T_ms = Tk_ms = 0.0
    launches = 0
    for _ in range(K):
        if world > 1:
            dist.barrier()
        ms, msk, nl = one_step()
        T_ms += ms; Tk_ms += msk; launches += nl
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([T_ms, Tk_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    T_ms, Tk_ms = tt.tolist()
    inter_step = float(n) * (n + NBV)
    value = inter_step * K / (T_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ---------------
    xs, zs, vxs, vzs = (np.zeros(n) for _ in range(4))
    dn = [cabi.dptr(a) for a in (xs, zs, vxs, vzs)]
    ke = max(5, min(K, 5))
    barrier()
    e0 = time.perf_counter()
    for _ in range(ke):
        cabi.check(L.unsflow_wake_upload(h, *up))
        cabi.check(L.unsflow_wake_step(h, 1, uinf, 0.0, dt))
        cabi.check(L.unsflow_wake_download(h, *dn))
    barrier()
    e2e_t = time.perf_counter() - e0
    te = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = inter_step * ke / te.item()
    extra = {}
    if world == 1:
        # the drop-in call itself: wakeroll(surf, curfield, dt) -> unsflow_wakeroll, host arrays
        xx, zz = x.copy(), z.copy()
        for rep in range(2):      # the first call allocates the per-thread device scratch, the second is steady state
            a0 = time.perf_counter()
            cabi.check(L.unsflow_wakeroll(n, cabi.dptr(xx), cabi.dptr(zz), up[2], up[3], dn[2], dn[3], NBV, up[4], up[5], up[6], up[7],
                                          uinf, 0.0, dt))
            extra["e2e_wakeroll_call_interactions_per_s"] = inter_step / (time.perf_counter() - a0)
        cabi.check(L.unsflow_release_scratch())
    # ---- verification on the same handle (all ranks take part): one step with dt = 0 on the
    # original field, 32 sampled targets against the oracle's long-double truth ----------------
    cabi.check(L.unsflow_wake_upload(h, *up))
    cabi.check(L.unsflow_wake_step(h, 1, 0.0, 0.0, 0.0))
    cabi.check(L.unsflow_wake_download(h, None, None, dn[2], dn[3]))
    verify = None
    if rank == 0:
        b_, e_ = C.c_int64(), C.c_int64()
        cabi.check(L.unsflow_partition(n, world, 0, 1, C.byref(b_), C.byref(e_)))
        hi = n if args.variant != 1 and n >= 16384 else e_.value    # direct: rank 0 only holds its slice
        idx = np.arange(0, hi, max(1, hi // 32))[:32]
        sx = np.concatenate([x, bx]); sz = np.concatenate([z, bz]); sg = np.concatenate([g, bs]); sv = np.concatenate([vc, bvc])
        tu, tw, au, aw = O.ind_vel_truth(sx, sz, sg, sv, x[idx], z[idx])
        verify = {"targets": int(len(idx)), "max_normalised_err": float(max(np.max(np.abs(vxs[idx] - tu)) / np.max(au),
                                                                             np.max(np.abs(vzs[idx] - tw)) / np.max(aw))),
                  "tolerance": 1e-12, "against": "oracle long-double sum"}
    # ---- extra: the multi-GPU split BASELINE.json names -- targets sliced contiguously over the ranks (one-sided
    # kernel, also the only path for non-uniform core radii), NCCL all-gather of the convected positions each step.
    # Timed for 2 steps; every rank verifies 8 targets of ITS OWN slice against the long-double truth.
    ag = None
    try:
        h1 = C.c_void_p()
        uid1 = (C.c_ubyte * 128)()
        if world > 1:
            t1 = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                cabi.check(L.unsflow_nccl_unique_id(uid1))
                t1.copy_(torch.tensor(list(uid1), dtype=torch.uint8))
            dist.broadcast(t1, 0)
            for i, v in enumerate(t1.cpu().tolist()):
                uid1[i] = v
        cabi.check(L.unsflow_wake_create(C.byref(h1), n, NBV, rank, world, uid1 if world > 1 else None, 1))
        cabi.check(L.unsflow_wake_upload(h1, *up))
        cabi.check(L.unsflow_wake_step(h1, 1, 0.0, 0.0, 0.0))          # dt = 0: velocities of the original field
        cabi.check(L.unsflow_wake_download(h1, None, None, dn[2], dn[3]))
        b1, e1 = C.c_int64(), C.c_int64()
        cabi.check(L.unsflow_partition(n, world, rank, 1, C.byref(b1), C.byref(e1)))
        idx1 = np.linspace(b1.value, e1.value - 1, 8).astype(np.int64)
        sx1 = np.concatenate([x, bx]); sz1 = np.concatenate([z, bz]); sg1 = np.concatenate([g, bs]); sv1 = np.concatenate([vc, bvc])
        tu1, tw1, au1, aw1 = O.ind_vel_truth(sx1, sz1, sg1, sv1, x[idx1], z[idx1])
        err1 = float(max(np.max(np.abs(vxs[idx1] - tu1)) / np.max(au1), np.max(np.abs(vzs[idx1] - tw1)) / np.max(aw1)))
        barrier()
        cabi.check(L.unsflow_wake_step(h1, 2, uinf, 0.0, dt))
        ms1, msk1, nl1 = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h1, C.byref(ms1), C.byref(msk1), C.byref(nl1)))
        # after the exchange every rank must hold the same positions: checksum of x over all ranks
        cabi.check(L.unsflow_wake_download(h1, dn[0], dn[1], None, None))
        cs = float(np.dot(xs, ((np.arange(n) % 97) + 1.0)))
        ta = torch.tensor([ms1.value, err1, cs, -cs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        msmax, errmax, csmax, ncsmin = ta.tolist()
        ag = {"split": "contiguous target slices + ncclAllGather of x, z (wake.cu, kernel variant 1 = one-sided)",
              "interactions_per_s": inter_step * 2 / (msmax * 1e-3), "ms_per_step": msmax / 2,
              "verify_max_normalised_err_over_ranks": errmax, "verify_targets_per_rank": 8, "tolerance": 1e-12,
              "positions_identical_on_all_ranks": bool(csmax == -ncsmin)}
        cabi.check(L.unsflow_wake_destroy(h1))
    except Exception as e:
        ag = {"error": str(e)[:200]}

    # ---- extra: opt-in FAST precision mode (bias-centred Newton rsqrt, <= 6.5e-13 per pair); the headline
    # above is FULL precision -- this block only documents what the relaxed mode would buy ---------------
    fast = None
    try:
        cabi.check(L.unsflow_set_precision(1))
        cabi.check(L.unsflow_wake_upload(h, *up))
        cabi.check(L.unsflow_wake_step(h, 1, 0.0, 0.0, 0.0))
        cabi.check(L.unsflow_wake_sync(h))
        cabi.check(L.unsflow_wake_step(h, 2, 0.0, 0.0, 0.0))
        ms_f, msk_f, nl_f = C.c_float(), C.c_float(), C.c_int64()
        cabi.check(L.unsflow_wake_timing(h, C.byref(ms_f), C.byref(msk_f), C.byref(nl_f)))
        cabi.check(L.unsflow_wake_download(h, None, None, dn[2], dn[3]))
        tf_ = torch.tensor([ms_f.value], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
        fast = {"interactions_per_s": inter_step * 2 / (tf_.item() * 1e-3), "precision": "fast (opt-in, not the headline)"}
        if rank == 0 and verify is not None:
            fast["max_normalised_err"] = float(max(np.max(np.abs(vxs[idx] - tu)) / np.max(au), np.max(np.abs(vzs[idx] - tw)) / np.max(aw)))
    except Exception as e:
        fast = {"error": str(e)[:200]}
    finally:
        L.unsflow_set_precision(0)
    cabi.check(L.unsflow_wake_destroy(h))

    # ---- extra: config C3, 4096 independent LDVM members x 500 steps, members sharded over
    # ranks (no communication); aggregate member-steps/s from the slowest rank's kernel time ----
    ens = None
    if args.ensemble_members > 0:
        try:
            import math
            nk = nl = int(round(math.sqrt(args.ensemble_members)))
            ens_steps = args.ensemble_steps
            tabs, lesp, base = [], [], None
            for k in np.linspace(0.05, 1.0, nk):
                kin = uf.KinemDef(uf.SinDef(0.0, 20.0 * math.pi / 180, float(k), 0.0), uf.ConstDef(0.0), uf.ConstDef(1.0))
                sf = uf.TwoDSurf("FlatPlate", 0.25, kin, [0.1])
                base = base or sf
                tb = uf.kinematics_table(sf, uf.TwoDFlowField(), ens_steps, 0.015)
                for lc in np.linspace(0.05, 0.35, nl):
                    tabs.append(tb); lesp.append(float(lc))
            mine = uf.ensemble_shard(np.array(tabs), lesp, rank, world)      # cost-balanced deal of the members over the ranks
            mats, ntev, nlev, ems = uf.ldvm_batch(base, np.array(tabs)[mine], np.array(lesp)[mine], 0.015)
            tm = torch.tensor([ems], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ens = {"members": nk * nl, "steps": ens_steps, "kernel_ms_max_over_ranks": tm.item(),
                   "member_steps_per_s": nk * nl * ens_steps / (tm.item() * 1e-3), "finite": bool(np.all(np.isfinite(mats))),
                   "workload": "C3: 64 reduced frequencies x 64 LESPcrit, SinDef pitch 20 deg, one CTA per member"}
        except Exception as e:
            ens = {"error": str(e)[:200]}

    # ---- extra: device-resident LDVM stepper (configs C2, C5, C1).  With several GPUs the roll-up pair tiles are
    # split over the ranks (unsflow_ldvm_set_parallel), everything else is replicated bit-identically. -----------
    if args.ldvm_steps > 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from cases import c1_surf, c2_surf, c5_2dof
        from unsflow_b200.solvers import LdvmHandle, _opts
        import hashlib

        def new_uid():
            u = (C.c_ubyte * 128)()
            t2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                cabi.check(L.unsflow_nccl_unique_id(u))
                t2.copy_(torch.tensor(list(u), dtype=torch.uint8))
            dist.broadcast(t2, 0)
            for i, v in enumerate(t2.cpu().tolist()):
                u[i] = v
            return u

        def stepper_run(surf, fld, driver, ns, dtstar, nranks, delv=None, twodof=None, warm=3):
            """ns timed steps after `warm` untimed ones; returns (device ms max over ranks, launches, mat, handle-final field)"""
            hh = LdvmHandle(surf, fld, _opts(driver, dtstar, ns + warm, delv or uf.delNone(), None))
            try:
                if nranks > 1:
                    hh.set_parallel(rank, nranks, new_uid())
                    extra["ldvm_exchange"] = "NVLink peer memory (fused finish + exchange kernel)" if hh.parallel_info()[2] else "NCCL all-reduce"
                if twodof:
                    hh.set_2dof(*twodof)
                hh.set_kinematics(uf.kinematics_table(surf, fld, ns + warm, dtstar, two_dof=twodof is not None))
                hh.run(warm)
                if nranks > 1:
                    barrier()
                mat = hh.run(ns)
                ms, nl = hh.timing()
                hh.download()
                cnt = hh.counts()
            finally:
                hh.close()
            tl = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if nranks > 1:
                dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            return tl.item(), nl, mat, cnt

        try:
            # config C2 (pitch-plunge, no deletion), windows of the growth to 1e5 vortices: wake pre-seeded with S-wake(N)
            win = {}
            for nw, nsw in ((1_000, 64), (10_000, 64), (100_000, args.ldvm_steps)):
                surf, fld, dtstar = c2_surf(uf)
                wx, wz, wg, wvc = s_wake(nw)
                fld.tev = uf.VortList.from_arrays(wx + 0.5, wz, wg, wvc)
                use_ranks = world if nw >= 100_000 else 1      # small wakes: replicas only (launch-bound)
                if use_ranks == 1 and rank != 0:
                    continue
                ms, nl, mat, cnt = stepper_run(surf, fld, cabi.DRIVER_LDVM, nsw, dtstar, use_ranks)
                win[str(nw)] = {"steps_per_s": nsw / (ms * 1e-3), "steps": nsw, "launches_per_step": nl / nsw, "ranks": use_ranks}
                if nw == 100_000:
                    extra["ldvm_steps_per_s_at_N1e5"] = nsw / (ms * 1e-3)
                    extra["ldvm_launches_per_step"] = nl / nsw
                    if world > 1:
                        # parity of the N-rank stepper in this very run: (i) every rank must return the same bits,
                        # (ii) rank 0 repeats the steps on ONE rank from the same state and reports the deviation
                        dig = hashlib.sha256(np.ascontiguousarray(mat).tobytes() + np.ascontiguousarray(fld.tev.x).tobytes()
                                             + np.ascontiguousarray(fld.tev.vz).tobytes()).digest()
                        td = torch.tensor(list(dig), dtype=torch.uint8, device="cuda")
                        allg = [torch.zeros_like(td) for _ in range(world)]
                        dist.all_gather(allg, td)
                        same = all(bool(torch.equal(allg[0], a)) for a in allg)
                        par = {"ranks_bit_identical": same, "steps": nsw + 3}
                        if rank == 0:
                            s1, f1, _ = c2_surf(uf)
                            f1.tev = uf.VortList.from_arrays(wx + 0.5, wz, wg, wvc)
                            ms1, nl1, mat1, cnt1 = stepper_run(s1, f1, cabi.DRIVER_LDVM, nsw, dtstar, 1)
                            par["max_abs_dmat_vs_1_rank"] = float(np.max(np.abs(mat - mat1)))
                            par["max_abs_dx_vs_1_rank"] = float(np.max(np.abs(fld.tev.x - f1.tev.x)))
                            vs = float(np.max(np.abs(f1.tev.vz)))
                            par["max_rel_dvz_vs_1_rank"] = float(np.max(np.abs(fld.tev.vz - f1.tev.vz)) / vs)
                            par["one_rank_steps_per_s"] = nsw / (ms1 * 1e-3)
                        extra["ldvm_multi_rank_parity"] = par
            if rank == 0:
                extra["ldvm_C2_windows"] = win
        except Exception as e:   # never lose the headline line
            extra["ldvm_error"] = str(e)[:300]
        try:
            # config C5: ldvm2DOF + delSpalart on a 1e5-vortex wake (merging active from step 1)
            surf, fld, strpar, kin0 = c5_2dof(uf)
            nw = 100_000
            wx, wz, wg, wvc = s_wake(nw)
            fld.tev = uf.VortList.from_arrays(wx + 0.5, wz, wg, wvc)
            ns5 = max(args.ldvm_steps, 40)
            ms, nl, mat, cnt = stepper_run(surf, fld, cabi.DRIVER_LDVM2DOF, ns5, 0.015, world, delv=uf.delSpalart(nw, 10.0, 1e-6),
                                           twodof=(strpar, kin0))
            if rank == 0:
                extra["ldvm2dof_C5"] = {"steps_per_s": ns5 / (ms * 1e-3), "steps": ns5, "wake0": nw, "tev_after": int(cnt[0]), "lev_after": int(cnt[1]),
                                        "merged_tevs": int(nw + ns5 + 3 - cnt[0]), "finite": bool(np.all(np.isfinite(mat))),
                                        "ranks": world, "launches_per_step": nl / ns5,
                                        "parity": "tests/test_gpu_ldvm.py::test_c5_2dof_spalart_merging_on_a_large_wake_against_the_oracle (5.5e4 wake, every merge decision vs the oracle)"}
        except Exception as e:
            extra["ldvm2dof_C5"] = {"error": str(e)[:300]}
        # config C1 (the reference's own test case, test/ldvmLinTest.jl through ldvm, 468 steps from an empty
        # wake): device-resident steps/s on rank 0, with the CPU oracle's steps/s beside it
        if rank == 0:
            try:
                surf1, fld1, n1, dts1 = c1_surf(uf)
                uf.ldvm(copy.deepcopy(surf1), uf.TwoDFlowField(), n1, dts1, write_summary=False)        # warm-up (graph capture)
                hh = LdvmHandle(surf1, fld1, _opts(cabi.DRIVER_LDVM, dts1 * surf1.c / surf1.uref, n1, uf.delNone(), None))
                hh.set_kinematics(uf.kinematics_table(surf1, fld1, n1, dts1))
                hh.run(n1)
                ms1, nl1 = hh.timing()
                hh.close()
                c1 = {"steps": n1, "device_ms": ms1, "steps_per_s": n1 / (ms1 * 1e-3), "us_per_step": 1e3 * ms1 / n1, "launches_per_step": nl1 / n1,
                      "path": "k_cluster_steps (64 steps per launch of one thread-block cluster)" if nl1 < n1 else "five kernels per step, CUDA graph"}
                if not args.no_cpu_baseline:
                    t0 = time.perf_counter()
                    O.Sim(surf1, uf.TwoDFlowField()).run(O.DRIVER_LDVM, uf.kinematics_table(surf1, uf.TwoDFlowField(), n1, dts1), dts1)
                    c1["cpu_oracle_steps_per_s"] = n1 / (time.perf_counter() - t0)
                extra["ldvm_C1"] = c1
            except Exception as e:
                extra["ldvm_C1"] = {"error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    ach = FLOP_PER_INTERACTION * inter_step * K / (Tk_ms * 1e-3) / 1e12 / 1.0   # per-rank kernel = 1/world of the pairs
    ach_per_gpu = ach / world
    line = {
        "metric": "FP64 vortex interactions/s", "value": value, "unit": "interactions/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": T_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 wakeroll step (calcs.jl:222-294), S-wake(N={n}) + {NBV} bound vortices, device-resident",
                   "n_vortices": n, "interactions_per_step": inter_step, "flop_per_interaction": FLOP_PER_INTERACTION,
                   "l2": "flushed between timed steps (256 MiB write); per-step partial sums exceed L2 anyway",
                   "kernel_variant": args.variant, "parallelism": f"pair work split over {world} rank(s), NCCL exchange per step"},
        "clocks": clocks, "wall_s_timed_region": wall,
        "e2e": {"value": e2e_val, "unit": "interactions/s", "h2d_bytes_per_step": int(8 * (4 * n + 4 * NBV)),
                "d2h_bytes_per_step": int(8 * 4 * n), "steps": ke,
                "how": "unsflow_wake_upload (host arrays) + unsflow_wake_step + unsflow_wake_download per step"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "achieved": ach_per_gpu, "peak": tf.value, "unit": "TFLOP/s",
                     "frac": ach_per_gpu / tf.value,
                     "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": ach_per_gpu / FP64_NOMINAL_TFLOPS,
                     "traffic": NCU_TRAFFIC_SYM8_N1E6 if (world == 1 and n == 1_000_000 and args.variant == 0) else None,
                     "kernel": "k_induce_* (all-pairs Vatistas induction)",
                     "peak_source": f"measured in this run: DFMA-chain microbenchmark (unsflow_fp64_peak), ~{mhz.value:.0f} MHz "
                                    "equivalent; MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                     "algorithmic": f"{FLOP_PER_INTERACTION:.0f} flop/interaction x {inter_step / world:.4e} interactions per launch per GPU"},
    }
    line["verify"] = verify
    line["allgather_target_slices"] = ag
    line["ensemble_C3"] = ens
    line["fast_precision_mode"] = fast
    line.update(extra)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_port(O)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
