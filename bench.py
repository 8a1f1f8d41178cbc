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
STAGE_W = {"inv_y_gemm": 3 + 5, "inv_x_pass": 5 + 6, "z_pass_nl": 6 + 3, "fwd_x_pass": 3 + 3, "fwd_y_gemm": 3 + 3, "tau_solve": 15 + 4}
STEP_W = 59


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gp = w["Nx"] * w["Ny"] * w["Nz"]
    Wbytes = 8 * w["Nx"] * w["Ny"] * 2 * (w["Nz"] // 2 + 1)
    nlname = {"rot": "rotational", "skew": "skew-symmetric"}.get(flags_kw(w)["nonlinearity"], flags_kw(w)["nonlinearity"])
    config = {"workload": "%s: SBDF3, %s NL, 2/3 dealiasing, FP64, dt=%g" % (w["desc"], nlname, w["dt"]), "grid": [w["Nx"], w["Ny"], w["Nz"]],
              "l2": "working set (>= 34 W = %.1f MB) %s the 126 MB L2" % (34 * Wbytes / 1e6, "exceeds" if 34 * Wbytes > 126e6 else "fits in")}

    if args.impl == "reference":
        if rank != 0:
            return
        ws = sample_grid(w)
        steps = max(1, min(args.steps, 3))
        ms = reference_steps(ws, steps, min(args.warmup, 1))
        val = ws["Nx"] * ws["Ny"] * ws["Nz"] / (ms * 1e-3)
        sample = "reference DNS (unmodified sources, FFT shim instead of FFTW), %dx%dx%d grid, %d SBDF3 steps" % (ws["Nx"], ws["Ny"], ws["Nz"], steps)
        print(json.dumps({"impl": "reference", "metric": "dns_grid_point_steps_per_s", "value": val, "unit": "grid-pt-steps/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms * gp / (ws["Nx"] * ws["Ny"] * ws["Nz"]),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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
    if world > 1:
        # one process per GPU: torch.distributed (NCCL) is the plumbing (id broadcast, barriers, max over ranks); the
        # data path's all-to-all / all-reduce are issued by libcfgpu.so on its own NCCL communicator
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(lib.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        lib.comm_init_nccl(rank, world, bytes(idt.cpu().numpy().tobytes()))
        config["parallelism"] = "kx-slab (spectral) / y-slab (physical) over %d GPUs, NCCL all-to-all" % world

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
    u0 = synthetic_field(w)
    ug = cf.FlowField(lib, w["Nx"], w["Ny"], w["Nz"], 3, w["Lx"], w["Lz"]).set(u0, padded=True)
    dns = cf.DNS(ug, cf.make_flags(**flags_kw(w)))
    dns.advance(2)            # SMRK2 initialisation steps of SBDF3 (not part of the metric)
    dns.advance(args.warmup)
    barrier()

    sampler = ClockSampler(dev)
    sampler.start()
    l0 = lib.launch_count()
    lib.profile_enable(True)
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
        n_box = n  # bytes that actually cross PCIe: de-aliased spectral fields travel as their retained box (own kx rows)
        Kx_, Kz_ = w["Nx"] // 3 - 1, w["Nz"] // 3 - 1
        x0_, x1_, _, _ = lib.comm_ranges(2 * Kx_ + 1, w["Ny"], rank)
        n_box = 3 * w["Ny"] * (x1_ - x0_) * (Kz_ + 1) * 2
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

    # ---- roofline of the dominant stage
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
    # y-GEMM stages are FP64 tensor-pipe (DMMA) work: flops of the even/odd-split contractions the kernels perform
    Kx_, Kz_ = w["Nx"] // 3 - 1, w["Nz"] // 3 - 1
    ncols = 2 * (2 * Kx_ + 1) * (Kz_ + 1) / world            # real columns of this rank
    gemm_flops = {"inv_y_gemm": 5, "fwd_y_gemm": 3}          # matrices applied: u,v,w + du/dy,dw/dy ; f_x,f_y,f_z
    for k_, nm in gemm_flops.items():
        if k_ in stages:
            fl = nm * 2.0 * w["Ny"] * ((w["Ny"] + 1) // 2) * ncols
            stages[k_]["TFLOPs"] = fl / (stages[k_]["ms_per_step"] * 1e-3) / 1e12
    dom = max((k for k in stages if k in STAGE_W), key=lambda k: stages[k]["ms_per_step"])
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if world == 1:
            traffic = tr.get(args.workload, {}).get(dom)
    except Exception:
        pass
    if dom in gemm_flops:
        # FP64 tensor work: MEASURED_PEAKS.json has no FP64 figure, so the yard-stick is cuBLAS DGEMM measured here
        def dgemm_peak(n=8192, reps=5):
            a = torch.randn(n, n, dtype=torch.float64, device="cuda")
            b = torch.randn(n, n, dtype=torch.float64, device="cuda")
            for _ in range(2):
                c = a @ b
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                c = a @ b  # noqa: F841
            e1.record()
            torch.cuda.synchronize()
            return 2.0 * n ** 3 / (e0.elapsed_time(e1) / reps * 1e-3) / 1e12
        try:
            f64_peak = dgemm_peak()
            src = "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 tensor figure)"
        except Exception:
            f64_peak, src = 40.0, "fallback: nominal B200 FP64 tensor rate"
        ach = stages[dom]["TFLOPs"]
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": f64_peak, "unit": "TFLOP/s", "frac": ach / f64_peak,
                    "traffic": traffic, "peak_source": src, "precision": "fp64 (DMMA m16n8k8)"}
    else:
        ach = stages[dom]["algorithmic_GBps"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}
    if traffic:  # what the kernel really moved (ncu), next to the algorithmic figure: SURVEY's byte counts charge full
        # padded arrays although only the retained 44 % of the modes are touched, and leave the solver's factors out
        roofline["dram_GBps"] = traffic / (stages[dom]["ms_per_step"] * 1e-3) / 1e9
        roofline["dram_frac"] = roofline["dram_GBps"] / hbm_peak
    roofline["hbm_peak_GBps"] = hbm_peak
    roofline["whole_step_algorithmic_GBps_per_gpu"] = STEP_W * Wbytes / world / (ms * 1e-3) / 1e9
    roofline["whole_step_frac"] = roofline["whole_step_algorithmic_GBps_per_gpu"] / hbm_peak
    for k_ in stages:  # every stage against its own bound, so the next kernel to work on can be read off the line
        if k_ in STAGE_W and k_ not in gemm_flops:
            stages[k_]["frac_of_hbm_peak"] = stages[k_]["algorithmic_GBps"] / hbm_peak

    if rank != 0:
        return
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        ws = sample_grid(w)
        cms = reference_steps(ws, 2, 0)
        cval = ws["Nx"] * ws["Ny"] * ws["Nz"] / (cms * 1e-3)
        cpu_baseline = {"value": cval, "unit": "grid-pt-steps/s", "cores": 1, "kind": "reference",
                        "sample": "reference DNS (unmodified sources + FFT shim), %dx%dx%d grid, 2 SBDF3 steps after 2 init steps" % (ws["Nx"], ws["Ny"], ws["Nz"])}

    print(json.dumps({"metric": "dns_grid_point_steps_per_s", "value": value, "unit": "grid-pt-steps/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                      "gpu_launches": launches, "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline}))


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
