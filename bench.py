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
          values are inside the timed region; e2e.copy_ceiling is the same byte traffic with NO kernel
          (the box's PCIe / host-memory ceiling, all ranks at once), e2e.pageable the same call on plain
          (unpinned) numpy arrays, which the library moves through its pinned staging ring
  roofline  FP64 tensor pipe: flops the kernel EXECUTES / CUDA-event kernel time, against the FP64 peak
          measured in this very run (MEASURED_PEAKS.json has no FP64 entry).  Plan-time group merging
          makes executed < the SURVEY §8(d) rule; the rule-based figure is reported beside it as
          algorithmic_achieved / algorithmic_frac
          Small-chi workloads (--config 6: chi = 1 product state, --config 7: chi = 2, both on config 2's layout and
          point set) run the table kernel and report an HBM roofline instead: algorithmic bytes (8 B per coordinate
          read + 8 / 16 B per value written) / kernel time against MEASURED_PEAKS.json's hbm_gbs
  configs   (default invocation, 1 GPU) short timed runs of the other BASELINE configurations — cfg3 (binary tree,
          chi = 64), cfg4 (the full 16384^2 grid + summed quadrature), cfg5 (complex chi = 128) — and the two
          small-chi HBM datapoints (cfg6, cfg7), each with its own value / kernel time / roofline / clocks
  fp64_peaks  the FP64 denominators measured in this run: DMMA and DFMA register loops, cuBLAS DGEMM burst
  multi_plan  (N > 1) rank 0 alone drives all N GPUs through ONE multi-device plan (ttn_plan_create_multi), the
          entry point a Julia caller of evaluate(f, pts; ngpus=N) uses
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
        return f, 4, 2 ** 21, ("cfg5: complex 2-D MPS, 40 vertices with a Real and an Imag binary index each "
                               "(physical dim 4), chi=128 complex, 2^21 random complex points/GPU per step "
                               "(BASELINE: 1e9 points over 8 GPUs; --points sets the per-GPU count)")
    if config == 3:
        g = t.named_binary_tree(7)
        ws = g.vertices()[7:]          # 120 of the 127 vertices carry a binary site index, the top 7 none
        s = t.continuous_siteinds(g, [ws[i::3] for i in range(3)])
        f = t.rand_itn(s, link_space=64, rng=20263, normalise=True)
        return f, 3, 2 ** 21, ("cfg3: 3-D binary tree depth 7 (127 vertices, 3x40 bits), chi=64, real, 2^21 random "
                               "points/GPU per step (BASELINE: 1e9 points; --points sets the count)")
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
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Index of the next sample: legs of one run share one sampler and slice its rows."""
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = self.rows[lo:hi]
        if not rows:  # a leg shorter than the sampling period: the neighbouring samples
            rows = self.rows[max(lo - 1, 0):(hi + 1 if hi is not None else None)]
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()


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


TRAFFIC_SOURCES = {
    2: "profiles/r01_cfg2_dmma_v6_final.txt", 6: "profiles/r01_table_chi1.txt", 7: "profiles/r01_table_chi2.txt",
}


def static_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from a committed
    `ncu --set full` capture (profiles/traffic.json) — hardware counters cannot be read inside an un-profiled
    run, so this is a recorded figure, labelled with its source, not a live measurement."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return None, None
    v = tj.get(f"cfg{config}")
    if v is None:
        return None, None
    src = tj.get(f"cfg{config}_source") or TRAFFIC_SOURCES.get(config)
    return v, f"ncu --set full capture {src} (recorded, per launch over the same point count; not re-measured in this run)"


def measure_fp64_peaks(local_rank, torch, _capi):
    """FP64 denominators measured in this run: DMMA and DFMA register loops (csrc/k_peak.cu)
    and a cuBLAS DGEMM burst (torch.matmul on float64 = cublasDgemm; library code, used only as a yardstick)."""
    dfma, dmma = C.c_double(), C.c_double()
    _capi.check(_capi.lib().ttn_measure_fp64_peak(local_rank, C.byref(dfma), C.byref(dmma)))
    n = 4096
    a = torch.rand((n, n), dtype=torch.float64, device=f"cuda:{local_rank}")
    b = torch.rand((n, n), dtype=torch.float64, device=f"cuda:{local_rank}")
    c = torch.empty((n, n), dtype=torch.float64, device=f"cuda:{local_rank}")
    for _ in range(3):
        torch.matmul(a, b, out=c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    dgemm = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    return {"dmma_m8n8k4_register_loop": dmma.value, "dfma_register_loop": dfma.value,
            "cublas_dgemm_4096_burst": dgemm, "unit": "TFLOP/s",
            "denominator": max(dfma.value, dmma.value),
            "note": "MEASURED_PEAKS.json has no FP64 figure; the roofline denominator is the larger register loop "
                    "(tcgen05 has no FP64 kind: DMMA.8x8x4 is the only FP64 tensor instruction on sm_100a)"}


class Buffers:
    """Device + pinned host arrays of one (points, columns, value width) shape; reused between configs."""

    def __init__(self, torch, local_rank, npts, ncol, nc_out, seed):
        dev = f"cuda:{local_rank}"
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.key = (npts, ncol, nc_out)
        self.x_dev = torch.rand((npts, ncol), dtype=torch.float64, device=dev, generator=gen)
        self.out_dev = torch.empty(npts * nc_out, dtype=torch.float64, device=dev)
        self.x_host = torch.empty((npts, ncol), dtype=torch.float64).pin_memory()
        self.x_host.copy_(self.x_dev)
        self.out_host = torch.empty(npts * nc_out, dtype=torch.float64).pin_memory()
        self.x_np, self.out_np = self.x_host.numpy(), self.out_host.numpy()
        if nc_out == 2:
            self.out_np = self.out_np.view(np.complex128)


def hbm_peak():
    try:
        return (float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]),
                "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth, read + write bytes)")
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def roofline_of(config, info, o_dev, npts, kernel_ms, peaks, _capi):
    flops_pp = info["flops_per_point"]
    exec_tf = o_dev.flops_executed / (kernel_ms * 1e-3) / 1e12
    achieved_tf = flops_pp * npts / (kernel_ms * 1e-3) / 1e12
    traffic, tsrc = static_traffic(config)
    if o_dev.kernel_used == _capi.TTN_KERNEL_TABLE:
        # small chi: a few flops per point against 8 B per coordinate + 8 / 16 B per value -> HBM roofline
        peak, src = hbm_peak()
        bpp = info["bytes_per_point"]
        gbs = bpp * npts / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes_per_point": bpp, "peak_source": src,
                "executed_flops_per_point": o_dev.flops_executed / npts, "algorithmic_flops_per_point": flops_pp,
                "fp64_tflops_executed": exec_tf,
                "note": "table kernel (k_chain_table.cu): groups of chain vertices pre-contracted at plan time into "
                        "shared-memory tables, one lookup per group; algorithmic bytes = 8 B per coordinate read + "
                        "8 (16 complex) B per value written"}
    peak_tf = peaks["denominator"]
    hbm_view = None
    if traffic:
        hp, _ = hbm_peak()
        hbm_view = {"dram_gb_per_s": traffic / (kernel_ms * 1e-3) / 1e9, "frac_of_hbm_peak": traffic / (kernel_ms * 1e-3) / 1e9 / hp,
                    "algorithmic_bytes": info["bytes_per_point"] * npts, "traffic_over_algorithmic": traffic / (info["bytes_per_point"] * npts),
                    "note": "recorded DRAM bytes of one launch / this run's kernel time.  Deep-table chains read two random "
                            "128-byte table rows per point (traffic >> algorithmic bytes by design: 5 DMMA rounds per point "
                            "instead of 13); neither pipe saturates — the kernel is bound by the latency of those gathers "
                            "(profiles/r02_cfg2_mma6_deep.txt)"}
    return {"bound": "tensor", "pipe": "FP64 tensor pipe (DMMA.8x8x4; tcgen05 has no FP64 kind)", "hbm_view": hbm_view,
            "achieved": exec_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": exec_tf / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": tsrc,
            "flops": "executed",
            "executed_flops_per_point": o_dev.flops_executed / npts,
            "algorithmic_flops_per_point": flops_pp,
            "algorithmic_achieved": achieved_tf,
            "algorithmic_frac": achieved_tf / peak_tf if peak_tf else None,
            "peak_source": "measured in this run (fp64_peaks): the larger of the DMMA m8n8k4 and DFMA register loops",
            "note": "achieved/frac count the flops the kernel EXECUTES (SURVEY 8(d) asks for that wherever an "
                    "algebraic saving makes executed < rule): plan-time contraction (merged chain positions, deep "
                    f"leaf/root tables, subtree tables) makes a point cost {o_dev.flops_executed / npts:.0f} flop "
                    f"instead of the rule's {flops_pp:.0f}.  algorithmic_achieved/algorithmic_frac use the rule's "
                    "flops / kernel time (the contract's literal definition) and can exceed the pipe's peak for that reason"}


class Ctx:
    pass


def measure(ctx, config, steps, warmup, points=0, bufs=None, extras=False):
    """One workload on this rank's GPU: device-resident legs, end-to-end legs, roofline.  Returns the JSON line as a
    dict (rank 0) — max over ranks already taken."""
    torch, dist, _capi = ctx.torch, ctx.dist, ctx._capi
    f, ncol, npts, desc = build_workload(config)
    if points:
        npts = int(points)
    plan = f.plan(device=ctx.local_rank)
    info = plan.info()
    nc_out = 2 if info["is_complex"] else 1
    if bufs is None or bufs.key != (npts, ncol, nc_out):
        bufs = Buffers(torch, ctx.local_rank, npts, ncol, nc_out, 1234 + ctx.rank)
    b = bufs

    def barrier():
        torch.cuda.synchronize()
        if ctx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, k=steps, w=warmup):
        for _ in range(w):
            step_fn()
        barrier()
        m0 = ctx.sampler.mark() if ctx.sampler else 0
        t0 = time.perf_counter()
        kms, launches, o = 0.0, 0, None
        for _ in range(k):
            o = step_fn()
            if o is not None:
                kms += o.kernel_ms
                launches += o.n_launches
        barrier()
        dt = time.perf_counter() - t0
        m1 = ctx.sampler.mark() if ctx.sampler else 0
        if ctx.world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{ctx.local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)  # max over ranks
            dt = float(tt.item())
        clocks = ctx.sampler.summary(m0, m1) if ctx.sampler else None
        return dt / k, kms / k, launches, o, clocks

    dt_dev, kernel_ms, launches, o_dev, clocks_dev = timed(
        lambda: plan.evaluate_device(b.x_dev.data_ptr(), npts, b.out_dev.data_ptr()))
    dt_e2e, _, _, o_e2e, clocks_e2e = timed(lambda: plan.evaluate_host(b.x_np, out=b.out_np)[1])
    # correctness guard inside the bench: device path == host path bit for bit on a slice
    # (host-buffer calls of deep-table chains that move every coordinate as a double run the image without deep tables —
    # PCIe-bound, k_chain_mma.cu "Light variant" — so the two paths may then differ by the rounding of two FP64 evaluation
    # orders; hybrid-quantised pinned calls run the deep image: bitwise equal)
    nchk = min(npts, 1 << 16) * nc_out
    dv, hv = b.out_dev[:nchk].cpu().numpy(), b.out_host[:nchk].numpy()
    same = bool((dv == hv).all())
    rel = float((np.abs(dv - hv) / np.maximum(np.abs(hv), 1e-3 * np.sqrt(np.mean(hv ** 2)) + 1e-300)).max())

    h2d, d2h = int(npts * ncol * 8), int(npts * nc_out * 8)   # the caller's arrays: float64 coordinates in, values out
    e2e = {"value": npts * ctx.world / dt_e2e, "unit": "points/s",
           "h2d_bytes_per_step": int(o_e2e.h2d_bytes) * ctx.world, "d2h_bytes_per_step": int(o_e2e.d2h_bytes) * ctx.world,
           "host_array_bytes_per_step": {"coords_f64": h2d * ctx.world, "values": d2h * ctx.world},
           "coords_quantised_on_host": bool(o_e2e.staged & 4),
           "ms_per_step": dt_e2e * 1e3, "clocks": clocks_e2e,
           "api": "ttn_evaluate(host pinned float64 buffers) via the Python mirror's Plan.evaluate_host; h2d / d2h bytes are what "
                  "the call really moved over PCIe (ttn_opts.h2d_bytes / d2h_bytes): where every coordinate's digits are the bits "
                  "of floor(x 2^L) the library's host threads quantise coordinates to that 32-bit grid index on the host "
                  "(bit-exact, 4 bytes per coordinate instead of 8) — every chunk of a pageable array, every second chunk of "
                  "a pinned one when this GPU has the host to itself (the copy engines read the chunks in between in place); "
                  "with several ranks per host (LOCAL_WORLD_SIZE > 1) pinned arrays travel as doubles"}
    if extras:
        # the same bytes with NO kernel: H2D of the coordinates and D2H of the values on two streams, all ranks at
        # once — what this box's PCIe / host memory system can move; e2e is reported as a fraction of it
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_only():
            with torch.cuda.stream(s1):
                b.x_dev.copy_(b.x_host, non_blocking=True)
            with torch.cuda.stream(s2):
                b.out_host.copy_(b.out_dev, non_blocking=True)
            s1.synchronize()
            s2.synchronize()

        dt_copy = timed(copy_only, k=max(3, min(steps, 5)), w=1)[0]
        e2e["copy_ceiling"] = {"ms_per_step": dt_copy * 1e3, "points_per_s": npts * ctx.world / dt_copy,
                               "gb_per_s": (h2d + d2h) * ctx.world / dt_copy / 1e9,
                               "what": "pinned H2D of the step's coordinates + D2H of its values, concurrently, no "
                                       "kernels, every rank at once (max over ranks)"}
        e2e["frac_of_copy_ceiling"] = dt_copy / dt_e2e   # > 1 where host-side quantisation moved fewer bytes than the float64 arrays hold
        # plain (pageable) numpy arrays: the library stages them through its pinned ring with host threads
        xp, op = np.array(b.x_np[: min(npts, 40_000_000)]), np.empty(min(npts, 40_000_000) * nc_out)
        op = op.view(np.complex128) if nc_out == 2 else op
        dt_page, _, _, o_pg, _ = timed(lambda: plan.evaluate_host(xp, out=op)[1], k=3, w=1)
        e2e["pageable"] = {"value": len(xp) * ctx.world / dt_page, "unit": "points/s", "points_per_step": len(xp),
                           "staged_bits": int(o_pg.staged),
                           "h2d_bytes_per_point": float(o_pg.h2d_bytes) / len(xp),
                           "what": "same call on unpinned numpy arrays (what a Julia Matrix{Float64} is): moved through "
                                   "the library's pinned staging ring by its host thread pool; staged_bits 4 = the "
                                   "coordinates were quantised to their 32-bit grid index on the way (bit-exact)"}
        del xp, op

    in_gb, out_gb = h2d / 1e9, d2h / 1e9
    if in_gb + out_gb > 0.3:
        l2_note = f"inputs_larger_than_l2 ({in_gb:.1f} GB coords + {out_gb:.1f} GB values per step)"
    else:
        l2_note = (f"inputs {in_gb + out_gb:.2f} GB per step; the per-vertex message / state workspaces the step streams "
                   "through HBM (several GB) evict them between steps")
    line = {
        "metric": "points_per_sec_fp64", "value": npts * ctx.world / dt_dev, "unit": "points/s", "n_gpus": ctx.world,
        "steps": steps, "warmup": warmup, "ms_per_step": dt_dev * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "points_per_gpu": npts, "kernel": _capi.KERNEL_NAMES[o_dev.kernel_used],
                   "flops_per_point": info["flops_per_point"], "l2": l2_note,
                   "device_eq_host_bitwise": same, "device_vs_host_max_floored_rel": rel, "rank0_numa_node": ctx.numa},
        "kernel_ms_events": kernel_ms,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks_dev,
        "roofline": roofline_of(config, info, o_dev, npts, kernel_ms, ctx.peaks, _capi),
    }
    return line, plan, bufs


def measure_grid(ctx, steps, warmup):
    """BASELINE config 4 as literally stated: the 2-D chi=32 MPS on the full 16384^2 grid with the summed-grid
    quadrature.  No coordinate bytes exist (the grid is generated on the device); a step = the whole grid + sum."""
    torch, _capi = ctx.torch, ctx._capi
    f, _, _, _ = build_workload(4)
    plan = f.plan(device=ctx.local_rank)
    n = 2 ** 14
    out = torch.empty(n * n, dtype=torch.float64, device=f"cuda:{ctx.local_rank}")
    for _ in range(warmup):
        plan.evaluate_grid([2.0 ** -14] * 2, [n, n], reduce_sum=True, out_ptr=out.data_ptr())
    torch.cuda.synchronize()
    m0 = ctx.sampler.mark() if ctx.sampler else 0
    t0 = time.perf_counter()
    kms = 0.0
    for _ in range(steps):
        _, o = plan.evaluate_grid([2.0 ** -14] * 2, [n, n], reduce_sum=True, out_ptr=out.data_ptr())
        kms += o.kernel_ms
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    m1 = ctx.sampler.mark() if ctx.sampler else 0
    kms /= steps
    # end to end: values D2H into a pinned host array every step (2.1 GB), sum returned
    oh = torch.empty(n * n, dtype=torch.float64).pin_memory()
    ohn = oh.numpy()
    g = _capi.ttn_grid()
    steps_arr, counts_arr = (C.c_double * 2)(2.0 ** -14, 2.0 ** -14), (C.c_int64 * 2)(n, n)
    g.n_coords, g.first, g.npts = 2, 0, n * n
    g.step, g.count = C.cast(steps_arr, C.POINTER(C.c_double)), C.cast(counts_arr, C.POINTER(C.c_int64))

    def e2e_step():
        oo = plan._opts("auto", True)
        _capi.check(_capi.lib().ttn_evaluate_grid(plan._h, C.byref(g), ohn.ctypes.data_as(C.c_void_p), C.byref(oo)))
        return oo

    e2e_step()
    t0 = time.perf_counter()
    for _ in range(max(2, steps // 2)):
        oo = e2e_step()
    dt_e2e = (time.perf_counter() - t0) / max(2, steps // 2)
    same = bool((out[: 1 << 16].cpu().numpy() == ohn[: 1 << 16]).all()) and oo.sum_out[0] == o.sum_out[0]
    peak = ctx.peaks["denominator"]
    ach = o.flops_executed / (kms * 1e-3) / 1e12
    rule = plan.info()["flops_per_point"] * n * n
    del oh
    return {
        "metric": "points_per_sec_fp64", "value": n * n / dt, "unit": "points/s", "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4: 2-D interleaved MPS 28 sites, chi=32, FULL 16384^2 grid + summed quadrature "
                               "(values written to HBM and summed)", "kernel": _capi.KERNEL_NAMES[o.kernel_used],
                   "grid_sum": o.sum_out[0], "device_eq_host_bitwise": same},
        "kernel_ms_events": kms, "gpu_launches": o.n_launches * steps,
        "clocks": ctx.sampler.summary(m0, m1) if ctx.sampler else None,
        "e2e": {"value": n * n / dt_e2e, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * n * 8,
                "ms_per_step": dt_e2e * 1e3, "api": "ttn_evaluate_grid(host pinned out) — the grid is generated on the "
                                                      "device, only the 2^28 values cross PCIe"},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "traffic": None, "flops": "executed", "executed_flops": o.flops_executed,
                     "algorithmic_flops": rule, "algorithmic_achieved": rule / (kms * 1e-3) / 1e12,
                     "algorithmic_frac": rule / (kms * 1e-3) / 1e12 / peak,
                     "note": f"{o.flops_executed:.4g} flop EXECUTED per grid (prefix sharing: level-by-level expansion; "
                             f"the per-point flop rule counts {rule:.4g})"},
    }


def measure_multi_plan(ctx, plan, bufs, steps, warmup):
    """N > 1: rank 0 ALONE drives all N GPUs through one multi-device plan (ttn_plan_create_multi: contiguous point
    blocks, one host thread + three streams per GPU, values written straight into the caller's array) — the entry
    point behind evaluate(f, pts; ngpus=N).  The other ranks idle at a CPU (gloo) barrier so their GPUs are free."""
    torch, dist = ctx.torch, ctx.dist
    import itna_b200 as t
    res = None
    if ctx.rank == 0:
        try:
            npts = bufs.key[0]
            total = npts * ctx.world
            mp = t.Plan(plan.packed, devices=list(range(ctx.world)))
            xh = torch.empty((total, bufs.key[1]), dtype=torch.float64).pin_memory()
            oh = torch.empty(total * bufs.key[2], dtype=torch.float64).pin_memory()
            for g in range(ctx.world):
                xh[g * npts:(g + 1) * npts].copy_(bufs.x_host)
            xn, on = xh.numpy(), oh.numpy()
            if bufs.key[2] == 2:
                on = on.view(np.complex128)
            for _ in range(max(1, min(warmup, 2))):
                mp.evaluate_host(xn, out=on)
            k = max(2, min(steps, 5))
            t0 = time.perf_counter()
            kms = 0.0
            for _ in range(k):
                _, o = mp.evaluate_host(xn, out=on)
                kms += o.kernel_ms
            dt = (time.perf_counter() - t0) / k
            nchk = min(npts, 1 << 16)
            ref = plan.evaluate_host(bufs.x_np[:nchk])[0]       # single-GPU plan, host buffers, same points
            same = bool((on[:nchk] == ref).all()) and bool((on[(ctx.world - 1) * npts:(ctx.world - 1) * npts + nchk] == ref).all())
            res = {"e2e_value": total / dt, "unit": "points/s", "n_devices": int(o.n_devices_used), "ms_per_step": dt * 1e3,
                   "kernel_ms_max_over_devices": kms / k, "points_per_step": total, "blocks_eq_single_gpu_bitwise": same,
                   "api": "one process, ttn_plan_create_multi + ttn_evaluate(host pinned buffers): evaluate(f, pts; ngpus=N)"}
            mp.close()
            del xh, oh
        except Exception as e:  # report, never take the headline line down
            res = {"error": repr(e)}
    dist.barrier(group=ctx.cpu_group)
    return res


def check_nccl_sharded(ctx, config):
    """N > 1: the one-process-per-GPU driver with its REAL collectives (parallel.evaluate_sharded: contiguous blocks of
    one shared point set, NCCL all_gather of the values, NCCL all_reduce of the quadrature sum), checked on rank 0
    against a single-GPU evaluation of the same points."""
    from itna_b200.parallel import evaluate_sharded
    f, ncol, _, _ = build_workload(config)
    n = 4_000_003
    pts = np.random.default_rng(4321).random((n, ncol))     # the same array on every rank
    if f.plan(device=ctx.local_rank).packed.is_complex:     # ComplexIndexMap: points are complex numbers
        pts = pts[:, 0::2] + 1j * pts[:, 1::2]
    t0 = time.perf_counter()
    full = evaluate_sharded(f, pts, device=ctx.local_rank)
    total = evaluate_sharded(f, pts, device=ctx.local_rank, reduce="sum")
    dt = time.perf_counter() - t0
    res = None
    if ctx.rank == 0:
        one = f.plan(device=ctx.local_rank).evaluate_host(pts)[0]
        res = {"points": n, "ranks": ctx.world, "backend": ctx.dist.get_backend(),
               "gathered_values_eq_single_gpu_bitwise": bool((full == one).all()),
               "allreduce_sum_rel_err": float(abs(total - one.sum()) / np.abs(one).sum()), "seconds": dt}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2,
                    help="2 (default, BASELINE configs[1]), 3, 4, 5 or 1; 6 / 7 = the HBM-bound small-chi datapoints "
                         "(chi = 1 product state / chi = 2 on config 2's layout)")
    ap.add_argument("--points", type=float, default=0, help="override points per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true",
                    help="default invocation only: skip the short runs of configs 3, 4 (grid), 5, 6, 7")
    ap.add_argument("--grid", action="store_true",
                    help="config 4 only: evaluate the FULL 16384^2 grid with the summed quadrature (ttn_evaluate_grid, "
                         "prefix-shared kernel) instead of random points")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup  # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import itna_b200  # noqa: F401
    from itna_b200 import _capi

    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.torch, ctx.dist, ctx._capi = torch, dist, _capi
    ctx.rank, ctx.local_rank, ctx.world = rank, local_rank, world
    ctx.cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctx.cpu_group = dist.new_group(backend="gloo")
    ctx.numa = bind_to_gpu_numa_node(local_rank, torch) if world > 1 else None
    ctx.peaks = measure_fp64_peaks(local_rank, torch, _capi)
    ctx.sampler = ClockSampler(local_rank).start() if rank == 0 else None

    try:
        if args.grid:
            assert args.config == 4 and world == 1, "--grid is a single-GPU, config-4 line"
            line = measure_grid(ctx, args.steps, args.warmup)
            line["fp64_peaks"] = ctx.peaks
            print(json.dumps(line), flush=True)
            return
        line, plan, bufs = measure(ctx, args.config, args.steps, args.warmup, points=args.points, extras=True)
        line["fp64_peaks"] = ctx.peaks
        if world > 1:
            mp = measure_multi_plan(ctx, plan, bufs, args.steps, args.warmup)
            try:
                sh = check_nccl_sharded(ctx, args.config)
            except Exception as e:
                sh = {"error": repr(e)}
            if rank == 0:
                line["multi_plan"] = mp
                line["nccl_sharded_check"] = sh
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            x_np = bufs.x_np
            r1 = cpu_baseline(plan.packed, x_np[:2000], 1)
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
        if world == 1 and args.config == 2 and not args.points and not args.no_side_configs:
            # the other BASELINE configurations and the HBM datapoints, short runs (driver-visible: VERDICT r1 #2)
            side = {}
            k, w = max(3, min(args.steps, 5)), 3
            for cfg in (6, 7, 3, 5):
                try:
                    sl, _, bufs = measure(ctx, cfg, k if cfg in (3, 5) else max(k, 10), w, bufs=bufs)
                    side[f"cfg{cfg}"] = sl
                except Exception as e:
                    side[f"cfg{cfg}"] = {"error": repr(e)}
                if cfg == 7:
                    plan = None
                    bufs = None  # release the 1e8-point arrays before the big plans
                    torch.cuda.empty_cache()
            try:
                side["cfg4_grid"] = measure_grid(ctx, k, w)
            except Exception as e:
                side["cfg4_grid"] = {"error": repr(e)}
            line["configs"] = side
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        if ctx.sampler:
            ctx.sampler.stop()
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
