#!/usr/bin/env python
"""bench.py — points/s of batched ITensorNetworkFunction evaluation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]

A "step" is one pass of the hot path (digits -> slice -> leaf-to-root contraction) over one batch
of synthetic points.  Default workload = BASELINE.json configs[1]: a 2-D function on the comb tree
named_comb_tree((2,30)) (dimension i on tooth i: a 60-vertex chain), chi = 16, real, evaluated at
10^8 uniformly random points per GPU (weak scaling: every rank evaluates its own 10^8 points with
the network replicated; there is no data-path collective, SURVEY §8 e).

  value   points/s with the coordinates already resident in HBM (device -> device)
  e2e     points/s through the C ABI with pinned HOST buffers: H2D of the coordinates and D2H of the
          values are inside the timed region
  roofline  FP64 tensor pipe: flops the kernel EXECUTES / CUDA-event kernel time, against the FP64 peak
          measured in this very run (MEASURED_PEAKS.json has no FP64 entry).  Plan-time group merging
          makes executed < the SURVEY §8(d) rule; the rule-based figure is reported beside it as
          algorithmic_achieved / algorithmic_frac
          Small-chi workloads (--config 6: chi = 1 product state, --config 7: chi = 2, both on config 2's layout and
          point set) run the table kernel and report an HBM roofline instead: algorithmic bytes (8 B per coordinate
          read + 8 / 16 B per value written) / kernel time against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline  the oracle's reference-style evaluation (two-way BP + exp(sum log), what
          scalar(alg="bp") does per point) timed on this box's host cores on a bounded sample

--impl reference times the same CPU restatement with all host threads (Julia is not installed in
this image and the reference's arithmetic is un-vendored third-party code, so the reference itself
cannot run).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def build_workload(config: int):
    import itna_b200 as t
    if config == 2:
        g = t.named_comb_tree((2, 30))
        s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
        f = t.rand_itn(s, link_space=16, rng=20262, normalise=True)
        return f, 2, 100_000_000, "cfg2: 2-D comb tree 2x30 bits (60-vertex chain), chi=16, real, 1e8 random points/GPU"
    if config == 4:
        s = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
        f = t.rand_itn(s, link_space=32, rng=20264, normalise=True)
        return f, 2, 2 ** 26, "cfg4 shape: 2-D interleaved MPS 28 sites, chi=32, real, 2^26 random points/GPU"
    if config == 5:
        s = t.complex_continuous_siteinds(t.named_grid((40, 1)), map_dimension=2)
        f = t.rand_itn(s, link_space=128, rng=20265, eltype=complex, normalise=True)
        return f, 4, 2 ** 21, ("cfg5 shape: complex 2-D MPS, 40 vertices with a Real and an Imag binary index each "
                               "(physical dim 4), chi=128 complex, 2^21 random complex points/GPU")
    if config == 3:
        g = t.named_binary_tree(7)
        ws = g.vertices()[7:]          # 120 of the 127 vertices carry a binary site index, the top 7 none
        s = t.continuous_siteinds(g, [ws[i::3] for i in range(3)])
        f = t.rand_itn(s, link_space=64, rng=20263, normalise=True)
        return f, 3, 2 ** 21, "cfg3 shape: 3-D binary tree depth 7 (127 vertices), chi=64, real, 2^21 random points/GPU"
    if config == 1:
        s = t.continuous_siteinds(t.named_grid((20, 1)))
        f = t.sin_itn(s, k=2.0, a=0.3, c=1.1)
        return f, 1, 10_000, "cfg1: 1-D sin QTT, 20-bit MPS, chi=2 complex, 1e4 points"
    if config in (6, 7):
        # the small-chi end of the north star ("add a chi=1 exp_itn product-state case so the
        # small-chi -> HBM GB/s clause has a datapoint"): config 2's layout and point set at chi = 1 / chi = 2
        g = t.named_comb_tree((2, 30))
        s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
        if config == 6:
            f = t.exp_itn(s, k=0.9, a=0.1, c=1.2, dim=1)
            return f, 2, 100_000_000, ("hbm datapoint: exp_itn product state (chi=1) on the 2-D comb tree 2x30 bits, real, "
                                       "1e8 random points/GPU")
        f = t.rand_itn(s, link_space=2, rng=20267, normalise=True)
        return f, 2, 100_000_000, "small-chi datapoint: random chi=2 chain on the 2-D comb tree 2x30 bits, real, 1e8 random points/GPU"
    raise SystemExit(f"config {config} is a parity-test case, not a bench line")


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(packed, coords_sample, threads):
    """Reference-style CPU evaluation (oracle ORACLE_BP) on a bounded sample.  Test-infrastructure
    code used here only as the timed baseline, never on the product path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    t0 = time.perf_counter()
    orc.evaluate(packed, coords_sample, orc.ORACLE_BP, nthreads=threads)
    return len(coords_sample) / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference algorithm, all host threads."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    import itna_b200 as t
    f, ncol, npts, desc = build_workload(args.config)
    packed = t.pack(f)
    threads = len(os.sched_getaffinity(0))  # all host cores (torchrun pins OMP_NUM_THREADS=1; ignore it)
    rng = np.random.default_rng(1)
    probe = rng.random((2000, ncol))
    rate = cpu_baseline(packed, probe, threads)
    per_step = min(4.0, 60.0 / max(args.steps + args.warmup, 1))  # whole run bounded to about a minute
    sample = int(max(2000, min(npts, rate * per_step)))
    pts = rng.random((sample, ncol))
    for _ in range(args.warmup):
        orc.evaluate(packed, pts, orc.ORACLE_BP, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.evaluate(packed, pts, orc.ORACLE_BP, nthreads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    line = {
        "impl": "reference", "metric": "points_per_sec_fp64", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample_points_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} points/step of the same workload; C restatement of evaluate() "
                                   "(greedy digits, slice, two-way BP, exp(sum log)), OpenMP over points; "
                                   "Julia reference not runnable in this image"},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank, torch):
    """Multi-rank runs: pin this process to the CPUs of the NUMA node its GPU hangs off, so that the
    pinned staging buffers (first touch) and the copy-engine traffic stay on the local socket.
    Best effort: returns the node number, or None when sysfs does not say."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except (OSError, ValueError, AttributeError):
        pass
    return None


def bench_grid(args, plan, rank, world, local_rank, torch):
    """BASELINE config 4 as literally stated: the 2-D chi=32 MPS on the full 16384^2 grid with the summed-grid
    quadrature.  No coordinate bytes exist (the grid is generated on the device); a step = the whole grid + sum."""
    import ctypes as C
    from itna_b200 import _capi
    assert args.config == 4 and world == 1, "--grid is a single-GPU, config-4 line"
    n = 2 ** 14
    dfma, dmma = C.c_double(), C.c_double()
    _capi.check(_capi.lib().ttn_measure_fp64_peak(local_rank, C.byref(dfma), C.byref(dmma)))
    out = torch.empty(n * n, dtype=torch.float64, device=f"cuda:{local_rank}")
    for _ in range(args.warmup):
        plan.evaluate_grid([2.0 ** -14] * 2, [n, n], reduce_sum=True, out_ptr=out.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kms = 0.0
    for _ in range(args.steps):
        _, o = plan.evaluate_grid([2.0 ** -14] * 2, [n, n], reduce_sum=True, out_ptr=out.data_ptr())
        kms += o.kernel_ms
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    kms /= args.steps
    peak = max(dfma.value, dmma.value)
    ach = o.flops_executed / (kms * 1e-3) / 1e12
    print(json.dumps({
        "metric": "points_per_sec_fp64", "value": n * n / dt, "unit": "points/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4: 2-D interleaved MPS 28 sites, chi=32, FULL 16384^2 grid + summed quadrature "
                               "(values written to HBM and summed)", "kernel": _capi.KERNEL_NAMES[o.kernel_used],
                   "grid_sum": o.sum_out[0]},
        "kernel_ms_events": kms, "gpu_launches": o.n_launches * args.steps,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "traffic": None,
                     "algorithmic": f"{o.flops_executed:.4g} flop EXECUTED per grid (prefix sharing; the per-point flop "
                                    f"rule would count {plan.info()['flops_per_point'] * n * n:.4g})"},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2,
                    help="2 (default, BASELINE configs[1]), 3, 4, 5 or 1; 6 / 7 = the HBM-bound small-chi datapoints "
                         "(chi = 1 product state / chi = 2 on config 2's layout)")
    ap.add_argument("--points", type=float, default=0, help="override points per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grid", action="store_true",
                    help="config 4 only: evaluate the FULL 16384^2 grid with the summed quadrature (ttn_evaluate_grid, "
                         "prefix-shared kernel) instead of random points")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import itna_b200 as t
    from itna_b200 import _capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    numa = bind_to_gpu_numa_node(local_rank, torch) if world > 1 else None
    f, ncol, npts, desc = build_workload(args.config)
    if args.points:
        npts = int(args.points)
    plan = f.plan(device=local_rank)
    if args.grid:
        return bench_grid(args, plan, rank, world, local_rank, torch)
    info = plan.info()
    nc_out = 2 if info["is_complex"] else 1
    flops_pp = info["flops_per_point"]

    # FP64 denominators measured in this run
    dfma, dmma = C.c_double(), C.c_double()
    _capi.check(_capi.lib().ttn_measure_fp64_peak(local_rank, C.byref(dfma), C.byref(dmma)))

    gen = torch.Generator(device=f"cuda:{local_rank}")
    gen.manual_seed(1234 + rank)
    x_dev = torch.rand((npts, ncol), dtype=torch.float64, device=f"cuda:{local_rank}", generator=gen)
    out_dev = torch.empty(npts * nc_out, dtype=torch.float64, device=f"cuda:{local_rank}")
    x_host = torch.empty((npts, ncol), dtype=torch.float64).pin_memory()
    x_host.copy_(x_dev)
    out_host = torch.empty(npts * nc_out, dtype=torch.float64).pin_memory()
    x_np, out_np = x_host.numpy(), out_host.numpy()
    if nc_out == 2:
        out_np = out_np.view(np.complex128)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return plan.evaluate_device(x_dev.data_ptr(), npts, out_dev.data_ptr())

    def step_e2e():
        _, o = plan.evaluate_host(x_np, out=out_np)
        return o

    def timed(step_fn):
        for _ in range(args.warmup):
            step_fn()
        barrier()
        t0 = time.perf_counter()
        kms, launches = 0.0, 0
        for _ in range(args.steps):
            o = step_fn()
            kms += o.kernel_ms
            launches += o.n_launches
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)  # max over ranks
            dt = float(tt.item())
        return dt, kms, launches, o

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    dt_dev, kms_dev, launches, o_dev = timed(step_device)
    clocks = sampler.stop() if sampler else None
    dt_e2e, kms_e2e, _, o_e2e = timed(step_e2e)

    # correctness guard inside the bench: device path == host path bit for bit on a slice
    chk = slice(0, 1 << 16)
    same = bool((out_dev[: (1 << 16) * nc_out].cpu().numpy() == out_host[: (1 << 16) * nc_out].numpy()).all())
    del chk

    total_pts = npts * world
    value = total_pts * args.steps / dt_dev
    e2e = total_pts * args.steps / dt_e2e
    kernel_ms = kms_dev / args.steps
    achieved_tf = flops_pp * npts / (kernel_ms * 1e-3) / 1e12
    in_gb, out_gb = npts * ncol * 8 / 1e9, npts * nc_out * 8 / 1e9
    if in_gb + out_gb > 0.3:
        l2_note = f"inputs_larger_than_l2 ({in_gb:.1f} GB coords + {out_gb:.1f} GB values per step)"
    else:
        l2_note = (f"inputs {in_gb + out_gb:.2f} GB per step; the per-vertex message / state workspaces the step streams "
                   "through HBM (several GB) evict them between steps")
    exec_tf = o_dev.flops_executed / (kernel_ms * 1e-3) / 1e12
    peak_tf = max(dfma.value, dmma.value)

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"cfg{args.config}")
            except (OSError, ValueError):
                traffic = None
        line = {
            "metric": "points_per_sec_fp64", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "points_per_gpu": npts, "kernel": _capi.KERNEL_NAMES[o_dev.kernel_used],
                       "flops_per_point": flops_pp, "l2": l2_note,
                       "device_eq_host_bitwise": same, "rank0_numa_node": numa},
            "kernel_ms_events": kernel_ms,
            "e2e": {"value": e2e, "unit": "points/s", "h2d_bytes_per_step": int(npts * ncol * 8) * world,
                    "d2h_bytes_per_step": int(npts * nc_out * 8) * world, "ms_per_step": dt_e2e / args.steps * 1e3,
                    "api": "ttn_evaluate(host pinned buffers) via the Python mirror's Plan.evaluate_host"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "pipe": "FP64 tensor pipe (DMMA.8x8x4; tcgen05 has no FP64 kind)",
                         "achieved": exec_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": exec_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "flops": "executed",
                         "executed_flops_per_point": o_dev.flops_executed / npts,
                         "algorithmic_flops_per_point": flops_pp,
                         "algorithmic_achieved": achieved_tf,
                         "algorithmic_frac": achieved_tf / peak_tf if peak_tf else None,
                         "peak_source": "measured in this run by ttn_measure_fp64_peak (MEASURED_PEAKS.json has no "
                                        f"FP64 figure): DMMA m8n8k4 register loop {dmma.value:.2f} TFLOP/s, DFMA "
                                        f"register loop {dfma.value:.2f} TFLOP/s; the larger is the denominator",
                         "note": "achieved/frac count the flops the kernel EXECUTES (SURVEY 8(d) asks for that "
                                 "wherever an algebraic saving makes executed < rule): the chain kernel pre-contracts "
                                 "groups of vertices at plan time (DESIGN.md 'Group merging'), so a point costs "
                                 f"{o_dev.flops_executed / npts:.0f} flop instead of the rule's {flops_pp:.0f}. "
                                 "algorithmic_achieved/algorithmic_frac use the rule's flops / kernel time "
                                 "(the contract's literal definition) and exceed the pipe's peak for that reason"},
        }
        if o_dev.kernel_used == _capi.TTN_KERNEL_TABLE:
            # small chi: a few flops per point against 8 B per coordinate + 8 / 16 B per value -> HBM roofline
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
            try:
                hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
                hbm_src = "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth, read + write bytes)"
            except (OSError, ValueError, KeyError):
                pass
            bpp = info["bytes_per_point"]
            gbs = bpp * npts / (kernel_ms * 1e-3) / 1e9
            line["roofline"] = {
                "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "traffic": traffic, "algorithmic_bytes_per_point": bpp, "peak_source": hbm_src,
                "executed_flops_per_point": o_dev.flops_executed / npts, "algorithmic_flops_per_point": flops_pp,
                "fp64_tflops_executed": exec_tf,
                "note": "table kernel (k_chain_table.cu): groups of chain vertices pre-contracted at plan time into "
                        "shared-memory tables, one lookup per group; algorithmic bytes = 8 B per coordinate read + "
                        "8 (16 complex) B per value written"}
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as orc
            threads = len(os.sched_getaffinity(0))
            xs = x_np[:2000]
            r1 = cpu_baseline(plan.packed, xs, 1)
            n1 = int(max(2000, min(len(x_np), r1 * 8)))
            r1 = cpu_baseline(plan.packed, x_np[:n1], 1)
            rN = cpu_baseline(plan.packed, x_np[:2000 * threads], threads)
            nN = int(max(2000, min(len(x_np), rN * 8)))
            rN = cpu_baseline(plan.packed, x_np[:nN], threads)
            line["cpu_baseline"] = {
                "value": rN, "unit": "points/s", "cores": threads, "kind": "port",
                "single_thread_value": r1,
                "sample": f"first {nN} points ({threads} OpenMP threads) / first {n1} points (1 thread) of the same "
                          "workload; C restatement of the reference's per-point evaluate (greedy digits, slice, "
                          "two-way BP + exp(sum log)); the Julia reference itself is single-threaded and not "
                          "runnable in this image"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
