#!/usr/bin/env python
"""bench.py -- DNS time-step throughput of the B200 hot path (and of the reference's CPU path beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c5|c2|c1|golden] [--impl reference]

A "step" is one SBDF3 time step (rotational nonlinearity, 2/3 dealiasing) of the workload grid on synthetic input
(divergence-free random perturbation of the laminar base flow, fixed seed).  One JSON line is printed by rank 0.

  value   grid-point-steps/s with the fields resident in HBM (CUDA events around K steps on the launching stream)
  e2e     the same metric through the public host API with HOST buffers: every step uploads the velocity field from
          pinned host memory, advances one step, and downloads the result (host<->device copies inside the timed region)
  roofline   dominant pipeline stage: algorithmic bytes (SURVEY.md 8(d), W = 8*Nx*Ny*2(Nz/2+1)) / measured stage time
  cpu_baseline  the reference's own C++ (oracle/_ref: unmodified sources + in-repo FFT shim), 1 host core, on a bounded
          sample grid with the same Ny (grid-point-steps/s is the size-independent unit)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs; Lx, Lz as in SURVEY.md 8(d)
    "c4": dict(Nx=512, Ny=257, Nz=512, Lx=4 * np.pi, Lz=2 * np.pi, nu=1.0 / 4000, dt=0.002, desc="turbulent-channel grid 512x257x512"),
    "c5": dict(Nx=256, Ny=129, Nz=256, Lx=4 * np.pi, Lz=2 * np.pi, nu=1.0 / 1000, dt=0.005, desc="256x129x256"),
    "c2": dict(Nx=128, Ny=97, Nz=128, Lx=2 * np.pi, Lz=np.pi, nu=1.0 / 1800, dt=0.01, desc="plane Poiseuille fixed flux 128x97x128",
               flags=dict(ulowerwall=0.0, uupperwall=0.0, constraint="bulkv", Ubulk=2.0 / 3, nonlinearity="skew")),
    "c1": dict(Nx=32, Ny=33, Nz=32, Lx=2 * np.pi, Lz=np.pi, nu=1.0 / 400, dt=0.02, desc="plane Couette 32x33x32"),
    "golden": dict(Nx=48, Ny=35, Nz=48, Lx=2 * np.pi / 1.14, Lz=2 * np.pi / 2.5, nu=1.0 / 400, dt=0.025, desc="48x35x48"),
}
STAGES = ["inv_y_gemm", "inv_x_pass", "z_pass_nl", "fwd_x_pass", "fwd_y_gemm", "tau_solve", "linear", "tau_setup", "slab_alltoall"]
# algorithmic bytes per stage in units of W (rotational SBDF-k with k=3)
# (the x/z passes move one field less than SURVEY's 61 W table: curl u is formed while the inverse x-pass loads, so
# 6 fields -- u and curl u -- instead of 7 go through it; 59 W per step)
# Algorithmic bytes per stage in units of W: SURVEY.md 8(d)'s tables, SBDF3 (k = 3).  Rotational form, 61 W per step of which
# the transforms + nonlinear term are 36 W (the kernels move one field less through the x/z passes -- curl u is formed
# while the inverse x-pass loads -- but the contract figure is SURVEY's).  Skew-symmetric (and the other non-rotational
# forms, same pipeline): 91 W, transforms + nonlinear term 72 W.
STAGE_W_ROT = {"inv_y_gemm": 3 + 5, "inv_x_pass": 5 + 7, "z_pass_nl": 7 + 3, "fwd_x_pass": 3 + 3, "fwd_y_gemm": 3 + 3, "tau_solve": 15 + 4}
STAGE_W_SKEW = {"inv_y_gemm": 3 + 6, "inv_x_pass": 6 + 9, "z_pass_nl": 9 + 9, "fwd_x_pass": 9 + 9, "fwd_y_gemm": 9 + 3, "tau_solve": 15 + 4}
TRANSFORM_STAGES = ("inv_y_gemm", "inv_x_pass", "z_pass_nl", "fwd_x_pass", "fwd_y_gemm")


def stage_table(nonlinearity):
    return STAGE_W_SKEW if nonlinearity in ("skew", "conv", "div", "alt") else STAGE_W_ROT


def synthetic_field(w, seed=1, magn=0.1):
    """Smooth divergence-free perturbation with no-slip walls, built directly in spectral space:
    u = curl(psi e_y)-like modes (u = dpsi/dz, w = -dpsi/dx, v = 0) times (1-y^2)^2 -> Chebyshev coefficients."""
    Nx, Ny, Nz = w["Nx"], w["Ny"], w["Nz"]
    Mz = Nz // 2 + 1
    rng = np.random.default_rng(seed)
    u = np.zeros((3, Ny, Nx, Mz), dtype=np.complex128)
    Kx, Kz = min(Nx // 3 - 1, 12), min(Nz // 3 - 1, 12)
    # g(y) = (1-y^2)^2 = 3/8 T0 - 1/2 T2 + 1/8 T4
    g = np.zeros(Ny); g[0], g[2], g[4] = 3.0 / 8, -0.5, 1.0 / 8
    for kx in range(-Kx, Kx + 1):
        for kz in range(0, Kz + 1):
            if kx == 0 and kz == 0:
                continue
            if kz == 0 and kx < 0:
                continue
            amp = (rng.standard_normal() + 1j * rng.standard_normal()) * 0.6 ** (abs(kx) + kz)
            a, c = 2 * np.pi * kx / w["Lx"], 2 * np.pi * kz / w["Lz"]
            mx = kx % Nx
            u[0, :, mx, kz] = 1j * c * amp * g
            u[2, :, mx, kz] = -1j * a * amp * g
            if kz == 0:  # keep the kz=0 plane Hermitian
                u[0, :, (-kx) % Nx, 0] = np.conj(u[0, :, mx, 0])
                u[2, :, (-kx) % Nx, 0] = np.conj(u[2, :, mx, 0])
    arr = u.view(np.float64)
    arr *= magn / max(np.sqrt(np.sum(np.abs(u) ** 2)), 1e-300)
    return arr


def make_input(cf, lib, w):
    """Initial field of the workload: the reference's `randomfield` rule (tools/randomfield.cpp:50-67: serial drand48
    sequence, seed 1, Gaussian coefficients with spectral decay 0.6 over the whole retained box, divergence-free, no-slip,
    rescaled to L2Norm 0.2) generated by this package's own FlowField::addPerturbations; falls back to the band-limited
    synthetic field when the host library predates it."""
    if hasattr(cf, "randomfield") and os.environ.get("CF_BENCH_INPUT", "randomfield") == "randomfield":
        try:
            return cf.randomfield(lib, w["Nx"], w["Ny"], w["Nz"], w["Lx"], w["Lz"], seed=1, magn=0.2, smooth=0.4).get()
        except AttributeError:
            pass
    return synthetic_field(w)


def flags_kw(w):
    kw = dict(nu=w["nu"], dt=w["dt"], ulowerwall=-1.0, uupperwall=1.0, baseflow="laminar", constraint="gradp",
              timestepping="sbdf3", initstepping="smrk2", nonlinearity="rot", dealiasing="xz")
    kw.update(w.get("flags", {}))
    return kw


class ClockSampler(threading.Thread):
    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.stop_flag = device, [], set(), False
        self.smax = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.smax = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def reference_steps(w_sample, steps, warmup):
    """Times the reference's own DNS (oracle/_ref) on the sample grid: 1 core. Returns (ms_per_step, info)."""
    from oracle import refcf
    from tests import parity  # noqa: F401
    rf = refcf.RefField(w_sample["Nx"], w_sample["Ny"], w_sample["Nz"], 3, w_sample["Lx"], w_sample["Lz"])
    rf.data[...] = synthetic_field(w_sample)
    rf.set_padded(True)
    dns = refcf.RefDNS(rf, refcf.make_flags(**flags_kw(w_sample)))
    dns.advance(2 + warmup)  # 2 SMRK2 initialisation steps of SBDF3 + warm-up
    t0 = time.perf_counter()
    dns.advance(steps)
    dt = time.perf_counter() - t0
    return 1e3 * dt / steps


def sample_grid(w):
    """Bounded CPU sample of the workload: same Ny (same per-mode solver cost), smaller Nx, Nz."""
    s = dict(w)
    while s["Nx"] * s["Ny"] * s["Nz"] > 3.5e6 and s["Nx"] > 32:
        s["Nx"] //= 2
        s["Nz"] //= 2
    return s


def box_rows(arr, Nx, Nz, x0, x1):
    """retained-box rows mxi in [x0, x1) (kx = 0..Kx, -Kx..-1) of a field array [Nd][Ny][Nx][2 Mz] as complex [Nd][Ny][rows][Kz+1]"""
    Kx, Kz = Nx // 3 - 1, Nz // 3 - 1
    c = arr.view(np.complex128)
    rows = [(m if m <= Kx else m - (2 * Kx + 1)) % Nx for m in range(x0, x1)]
    return c[:, :, rows, :Kz + 1]


def small_case_parity(cf, lib, rank, world):
    """The 48x49x32 plane-Couette case of tests/mp_slab_worker.py, 4 SBDF3 steps on all ranks, against the committed
    fixture tests/golden/mp_slab_48x49x32.npz (computed by the compiled reference, tests/golden/make_mp_fixture.py):
    over the three exchanges: peer-memory push kernels with device-side flags (the default), stores over NVLink
    fused into the transform kernels (CFGPU_PEER_MODE=fused), staged NCCL send/recv (CFGPU_NO_PEER=1).  Relative L2 error of the all-gathered field, CFL, L2Norm."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "mp_slab_48x49x32.npz"))
    Nx, Ny, Nz = int(fx["Nx"]), int(fx["Ny"]), int(fx["Nz"])
    Kx, Kz = Nx // 3 - 1, Nz // 3 - 1
    from tests import parity as _p  # C1 flags (no oracle call below: parity.py only holds the case definition)
    flags = dict(_p.C1["flags"])
    Mz = Nz // 2 + 1
    u0 = np.zeros((3, Ny, Nx, Mz), dtype=np.complex128)
    rows = [(m if m <= Kx else m - (2 * Kx + 1)) % Nx for m in range(2 * Kx + 1)]
    u0[:, :, rows, :Kz + 1] = fx["u0"]
    out = {}
    saved = {k: os.environ.get(k) for k in ("CFGPU_NO_PEER", "CFGPU_PEER_MODE")}
    for leg in ("peer_push", "peer_fused", "staged"):
        os.environ.pop("CFGPU_NO_PEER", None)
        os.environ["CFGPU_PEER_MODE"] = "push" if leg == "peer_push" else "fused"
        if leg == "staged":
            os.environ["CFGPU_NO_PEER"] = "1"
        try:
            ug = cf.FlowField(lib, Nx, Ny, Nz, 3, float(fx["Lx"]), float(fx["Lz"]), float(fx["a"]), float(fx["b"])).set(u0.view(np.float64), padded=True)
            dns = cf.DNS(ug, cf.make_flags(**flags))
            cfl0 = dns.cfl()
            dns.advance(int(fx["nsteps"]))
            u1, _ = dns.get()
            norm = u1.l2norm()
            cfl4 = dns.cfl()
            u1.allgather()
            u1.set_padded(False)  # full download: every rank holds the complete field after the gather
            got = box_rows(u1.get(), Nx, Nz, 0, 2 * Kx + 1)
            ref = fx["u4"]
            out[leg] = {"rel_l2": float(np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel())),
                        "cfl0_rel": abs(cfl0 - float(fx["cfl0"])) / float(fx["cfl0"]), "cfl_rel": abs(cfl4 - float(fx["cfl4"])) / float(fx["cfl4"]),
                        "norm_rel": abs(norm - float(fx["norm4"])) / float(fx["norm4"])}
            del dns, ug, u1
        finally:
            for k, v in saved.items():  # back to what the caller's environment said
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return out


def state_signature(lib, dns, Nx, Ny, Nz, rank):
    """L2Norm, CFL and a per-kx-row checksum (sum |u|^2 over components, y, kz) of the DNS state: this rank's rows."""
    u, _ = dns.get()
    norm, cfl = u.l2norm(), dns.cfl()
    Kx = Nx // 3 - 1
    x0, x1, _, _ = lib.comm_ranges(2 * Kx + 1, Ny, rank)
    rows = box_rows(u.get(), Nx, Nz, x0, x1)  # a padded download only fetches this rank's rows
    return norm, cfl, np.sum(np.abs(rows) ** 2, axis=(0, 1, 3)), (x0, x1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="weak: Nx of the workload grid grows with the number of GPUs (per-GPU work fixed)")
    ap.add_argument("--stepper", default=None, help="time stepping scheme (default: the workload's, sbdf3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity block (N > 1)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "weak" and world > 1:
        w["Nx"] *= world
        w["Lx"] *= world
        w["desc"] = "%dx%dx%d (weak scaling: Nx, Lx x %d)" % (w["Nx"], w["Ny"], w["Nz"], world)
    if args.stepper:
        w.setdefault("flags", {})
        w["flags"] = dict(w["flags"], timestepping=args.stepper)
    fkw = flags_kw(w)
    STAGE_W = stage_table(fkw["nonlinearity"])
    STEP_W = sum(STAGE_W.values())
    # "transforms + nonlinear term" = 36 W of the rotational step (66 W skew-symmetric): the forward y-pass bytes
    # are booked with the solve there.  The time below is nevertheless that of ALL FIVE transform stages (the round-1
    # review's convention: 36 W / (inv y + inv x + z + fwd x + fwd y)), so the fraction is a lower bound.
    TRANS_W = 36 if STAGE_W is STAGE_W_ROT else 66
    gp = w["Nx"] * w["Ny"] * w["Nz"]
    Wbytes = 8 * w["Nx"] * w["Ny"] * 2 * (w["Nz"] // 2 + 1)
    nlname = {"rot": "rotational", "skew": "skew-symmetric"}.get(fkw["nonlinearity"], fkw["nonlinearity"])
    config = {"workload": "%s: %s, %s NL, 2/3 dealiasing, FP64, dt=%g" % (w["desc"], fkw["timestepping"].upper(), nlname, w["dt"]),
              "grid": [w["Nx"], w["Ny"], w["Nz"]],
              "l2": "working set (>= 34 W = %.1f MB) %s the 126 MB L2" % (34 * Wbytes / 1e6, "exceeds" if 34 * Wbytes > 126e6 else "fits in")}

    if args.impl == "reference":
        if rank != 0:
            return
        ws = sample_grid(w)
        steps = max(1, min(args.steps, 3))
        ms = reference_steps(ws, steps, min(args.warmup, 1))
        val = ws["Nx"] * ws["Ny"] * ws["Nz"] / (ms * 1e-3)
        sample = "reference DNS (unmodified sources, FFT shim instead of FFTW), %dx%dx%d sample grid (same Ny, Nx and Nz reduced), %d %s steps, 1 host core" % (
            ws["Nx"], ws["Ny"], ws["Nz"], steps, fkw["timestepping"].upper())
        config["workload"] += " -- reference arm timed on a bounded %dx%dx%d sample of it; ms_per_step is per SAMPLE step, " \
                              "value (grid-pt-steps/s) is the size-independent unit" % (ws["Nx"], ws["Ny"], ws["Nz"])
        config["sample_grid"] = [ws["Nx"], ws["Ny"], ws["Nz"]]
        print(json.dumps({"impl": "reference", "metric": "dns_grid_point_steps_per_s", "value": val, "unit": "grid-pt-steps/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms,
                          "ms_per_step_extrapolated_to_workload": ms * gp / (ws["Nx"] * ws["Ny"] * ws["Nz"]),
                          "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config, "steps_per_s_at_workload": val / gp,
                          "cpu_baseline": {"value": val, "unit": "grid-pt-steps/s", "cores": 1, "kind": "reference", "sample": sample},
                          "e2e": {"value": val, "unit": "grid-pt-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import channelflow_b200 as cf
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("CFGPU_DEVICE", str(dev))
    lib = cf.HostLib()
    dist = None
    u0 = make_input(cf, lib, w)

    # ---- multi-GPU parity, part 1: the same steps on ONE GPU (every rank runs them on its own device before the
    # communicator exists), to be compared with the N-rank result after the timed region
    parity_block = None
    nsteps_total = 2 + args.warmup + args.steps
    sig1 = None
    if world > 1 and not args.no_parity and args.scaling == "strong":
        ug1 = cf.FlowField(lib, w["Nx"], w["Ny"], w["Nz"], 3, w["Lx"], w["Lz"]).set(u0, padded=True)
        dns1 = cf.DNS(ug1, cf.make_flags(**fkw))
        dns1.advance(nsteps_total)
        n1, c1, rows1, _ = state_signature(lib, dns1, w["Nx"], w["Ny"], w["Nz"], 0)
        sig1 = (n1, c1, rows1)
        del dns1, ug1

    if world > 1:
        # one process per GPU: torch.distributed (NCCL) is the plumbing (id broadcast, barriers, max over ranks); the
        # data path's exchange / all-reduce are issued by libcfgpu.so (peer-memory stores over NVLink, its own NCCL communicator)
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(lib.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        lib.comm_init_nccl(rank, world, bytes(idt.cpu().numpy().tobytes()))
        config["parallelism"] = "kx-slab (spectral) / y-slab (physical) over %d GPUs, all-to-all fused into the kernels' stores over NVLink peer memory" % world

    def barrier():
        lib.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world > 1 and not args.no_parity:
        parity_block = {"small_case_48x49x32_vs_reference": small_case_parity(cf, lib, rank, world), "tolerance": 1e-12}

    ug = cf.FlowField(lib, w["Nx"], w["Ny"], w["Nz"], 3, w["Lx"], w["Lz"]).set(u0, padded=True)
    dns = cf.DNS(ug, cf.make_flags(**fkw))
    dns.advance(2)            # SMRK2 initialisation steps of SBDF3 (not part of the metric)
    dns.advance(args.warmup)
    barrier()

    sampler = ClockSampler(dev)
    sampler.start()
    l0 = lib.launch_count()
    # launch-bound grids (<= 2^20 points) are advanced by CUDA-graph replay (host/dnsalgo.cpp), which the per-stage event
    # timers would disable: no stage table for them
    stage_profile = w["Nx"] * w["Ny"] * w["Nz"] > (1 << 20) or os.environ.get("CFGPU_GRAPH") == "0"
    lib.profile_enable(stage_profile)
    lib.profile_read(reset=True)
    lib.timer_start()
    dns.advance(args.steps)
    ms_total = lib.timer_stop()
    barrier()
    ms_total = max_over_ranks(ms_total)
    stage_ms, stage_calls = lib.profile_read(reset=True)
    lib.profile_enable(False)
    launches = lib.launch_count() - l0
    ms = ms_total / args.steps
    value = gp / (ms * 1e-3)

    # ---- multi-GPU parity, part 2: the N-rank state after the timed steps against the 1-rank run of the same steps
    if parity_block is not None and sig1 is not None:
        nN, cN, rowsN, (x0, x1) = state_signature(lib, dns, w["Nx"], w["Ny"], w["Nz"], rank)
        ref_rows = sig1[2][x0:x1]
        err_rows = float(np.max(np.abs(rowsN - ref_rows) / np.maximum(np.abs(ref_rows), 1e-300 + 1e-13 * np.max(sig1[2]))))
        err_rows = max_over_ranks(err_rows)
        parity_block["c4_state_vs_1rank_same_steps"] = {
            "steps": nsteps_total, "l2norm_rel": abs(nN - sig1[0]) / sig1[0], "cfl_rel": abs(cN - sig1[1]) / abs(sig1[1]),
            "kx_row_checksum_max_rel": err_rows, "kx_rows_checked": int(len(sig1[2]))}
    if parity_block is not None:
        sm = parity_block["small_case_48x49x32_vs_reference"]
        ok = all(v < 1e-12 for leg in sm.values() for v in leg.values())
        big = parity_block.get("c4_state_vs_1rank_same_steps")
        if big:
            ok = ok and big["l2norm_rel"] < 1e-11 and big["cfl_rel"] < 1e-11 and big["kx_row_checksum_max_rel"] < 1e-9
        parity_block["ok"] = bool(ok)

    # ---- end to end through the host API with host buffers
    e2e = None
    if not args.no_e2e:
        n = int(np.prod(ug.shape))
        pin_in = torch.empty(n, dtype=torch.float64).pin_memory()
        pin_out = torch.empty(n, dtype=torch.float64).pin_memory()
        hin, hout = pin_in.numpy(), pin_out.numpy()
        cur, _ = dns.get()
        hin[:] = cur.get().ravel()
        ke = max(2, min(args.steps, 5))
        import ctypes as C
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        Kx_, Kz_ = w["Nx"] // 3 - 1, w["Nz"] // 3 - 1
        x0_, x1_, _, _ = lib.comm_ranges(2 * Kx_ + 1, w["Ny"], rank)
        n_box = 3 * w["Ny"] * (x1_ - x0_) * (Kz_ + 1) * 2  # what crosses PCIe: the retained box of this rank's kx rows
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            lib.L.cf_field_upload(cur.h, dp(hin))      # H2D of this step's input (pinned)
            dns.set(cur)
            dns.advance(1)
            lib.L.cf_dns_get(dns.h, cur.h, None)
            lib.L.cf_field_download(cur.h, dp(hout))   # D2H of this step's result
            hin, hout = hout, hin
        barrier()
        ems = max_over_ranks(1e3 * (time.perf_counter() - t0) / ke)
        e2e = {"value": gp / (ems * 1e-3), "unit": "grid-pt-steps/s", "h2d_bytes_per_step": 8 * n_box, "d2h_bytes_per_step": 8 * n_box,
               "ms_per_step": ems, "steps": ke,
               "note": "per rank; the host FlowField arrays are full size, only the retained (de-aliased) modes of the rank's kx rows cross PCIe"}
    clocks = sampler.result()

    # ---- tau-solver setup (NSE::reset_lambda: once per dt change, every cfDSI evaluation): timed outside the step loop
    setup_info = None
    try:
        lib.profile_enable(True)
        lib.profile_read(reset=True)
        dns.reset_dt(w["dt"])
        sms, scalls = lib.profile_read(reset=True)
        lib.profile_enable(False)
        if scalls[7]:
            setup_info = {"ms_per_lambda": sms[7] / scalls[7], "lambdas": scalls[7]}
    except Exception:
        pass

    # ---- roofline
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    stages = {}
    for name, t, c in zip(STAGES, stage_ms, stage_calls):
        if c:
            per = t / args.steps
            stages[name] = {"ms_per_step": per, "calls_per_step": c / args.steps}
            if name in STAGE_W:
                stages[name]["algorithmic_GBps"] = STAGE_W[name] * Wbytes / world / (per * 1e-3) / 1e9  # this rank's share
    # The y stages (named *_y_gemm after their first implementation) run as shared-memory FFTs (csrc/yfft.cu, HBM-bound) when
    # 2(Ny-1) factors into 2, 3, 5 and CF_YFFT is not 0; otherwise as FP64 tensor-pipe (DMMA) contractions, for which the flops
    # of the even/odd-split contractions are reported against cuBLAS DGEMM.
    def _smooth(n):
        for r in (2, 3, 5):
            while n % r == 0:
                n //= r
        return n == 1
    y_as_fft = os.environ.get("CF_YFFT", "1") != "0" and w["Ny"] >= 3 and _smooth(2 * (w["Ny"] - 1))
    Kx_, Kz_ = w["Nx"] // 3 - 1, w["Nz"] // 3 - 1
    ncols = 2 * (2 * Kx_ + 1) * (Kz_ + 1) / world            # real columns of this rank
    rot = STAGE_W is STAGE_W_ROT
    gemm_flops = {} if y_as_fft else {"inv_y_gemm": 5 if rot else 6, "fwd_y_gemm": 3 if rot else 6}   # matrices applied per step
    for k_, nm in gemm_flops.items():
        if k_ in stages:
            fl = nm * 2.0 * w["Ny"] * ((w["Ny"] + 1) // 2) * ncols
            stages[k_]["TFLOPs"] = fl / (stages[k_]["ms_per_step"] * 1e-3) / 1e12
    for k_ in ("inv_y_gemm", "fwd_y_gemm"):
        if k_ in stages:
            stages[k_]["kernel"] = "yfft_half_kernel (shared-memory FFT, hbm-bound)" if y_as_fft else "ygemm_kernel (DMMA contraction, fp64-tensor-bound)"
    # The roofline object describes the quantity the target is stated on: transforms + nonlinear term (36 W of
    # the 61 W rotational step) against the HBM roofline; `kernel` names the slowest of those stages.  Every stage's own
    # figure is in `stages`.
    tstages = [k for k in TRANSFORM_STAGES if k in stages]
    if not tstages:
        # no per-stage timers (CUDA-graph replay of a launch-bound grid whose state lives in L2): the whole step stands in
        stages["whole_step"] = {"ms_per_step": ms, "calls_per_step": 1.0, "algorithmic_GBps": STEP_W * Wbytes / world / (ms * 1e-3) / 1e9,
                                "note": "CUDA-graph replay, no stage timers; L2-resident state: the HBM fraction is not meaningful"}
        tstages = ["whole_step"]
    t_trans = sum(stages[k]["ms_per_step"] for k in tstages)
    dom = max(tstages, key=lambda k: stages[k]["ms_per_step"])
    trans_GBps = (TRANS_W if tstages != ["whole_step"] else STEP_W) * Wbytes / world / (t_trans * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "transforms+nonlinear term (%s); slowest stage: %s" % ("+".join(tstages), dom),
                "achieved": trans_GBps, "peak": hbm_peak, "unit": "GB/s", "frac": trans_GBps / hbm_peak,
                "traffic": None,  # dram bytes come from ncu captures (profiles/r02_*), never from a run under the profiler
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes": "SURVEY.md 8(d): %d W of the %d W step, W = 8*Nx*Ny*2(Nz/2+1) = %.1f MB%s" % (
                    TRANS_W, STEP_W, Wbytes / 1e6, ", this rank's 1/%d share" % world if world > 1 else ""),
                "transforms_nl_ms": t_trans, "transforms_nl_frac": trans_GBps / hbm_peak,
                "slowest_stage": dom, "slowest_stage_GBps": stages[dom].get("algorithmic_GBps"),
                "hbm_peak_GBps": hbm_peak}
    roofline["whole_step_algorithmic_GBps_per_gpu"] = STEP_W * Wbytes / world / (ms * 1e-3) / 1e9
    roofline["whole_step_frac"] = roofline["whole_step_algorithmic_GBps_per_gpu"] / hbm_peak
    for k_ in stages:  # every stage against its own bound, so the next kernel to work on can be read off the line
        if k_ in STAGE_W and k_ not in gemm_flops:
            stages[k_]["frac_of_hbm_peak"] = stages[k_]["algorithmic_GBps"] / hbm_peak
    if world == 1 and not args.no_cpu_baseline and gemm_flops:
        # FP64 tensor yard-stick for the y-GEMM stages: cuBLAS DGEMM measured in this run (MEASURED_PEAKS.json has no FP64 figure)
        try:
            n_ = 8192
            a_ = torch.randn(n_, n_, dtype=torch.float64, device="cuda")
            b_ = torch.randn(n_, n_, dtype=torch.float64, device="cuda")
            for _ in range(2):
                c_ = a_ @ b_
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                c_ = a_ @ b_  # noqa: F841
            e1.record()
            torch.cuda.synchronize()
            f64_peak = 2.0 * n_ ** 3 / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e12
            roofline["fp64_tensor_yardstick_TFLOPs"] = f64_peak
            for k_ in gemm_flops:
                if k_ in stages:
                    stages[k_]["frac_of_cublas_dgemm"] = stages[k_]["TFLOPs"] / f64_peak
            del a_, b_, c_
        except Exception:
            pass

    if parity_block is not None and not parity_block["ok"]:
        if rank == 0:
            print(json.dumps({"metric": "dns_grid_point_steps_per_s", "error": "multi-GPU parity failed", "multi_gpu_parity": parity_block}))
        sys.exit(3)
    if rank != 0:
        return
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        ws = sample_grid(w)
        cms = reference_steps(ws, 2, 0)
        cval = ws["Nx"] * ws["Ny"] * ws["Nz"] / (cms * 1e-3)
        cpu_baseline = {"value": cval, "unit": "grid-pt-steps/s", "cores": 1, "kind": "reference",
                        "sample": "reference DNS (unmodified sources + FFT shim), %dx%dx%d grid, 2 %s steps after 2 init steps" % (
                            ws["Nx"], ws["Ny"], ws["Nz"], fkw["timestepping"].upper())}

    line = {"metric": "dns_grid_point_steps_per_s", "value": value, "unit": "grid-pt-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline}
    if parity_block is not None:
        line["multi_gpu_parity"] = parity_block
    if setup_info is not None:
        line["tau_setup"] = setup_info
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        try:
            import torch.distributed as _dist
            if _dist.is_available() and _dist.is_initialized():
                _dist.destroy_process_group()
        except Exception:
            pass
