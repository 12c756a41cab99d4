"""cfg2 (2x30 bits, chi = 16): device-resident rate and plan time vs the deep leaf / root table budget TTN_MMA_DEEP."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
import bench
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
x = torch.rand((npts, 2), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
for b in [int(a) for a in os.environ.get("SWEEP_B", "0,12,14,16,17,18,19,20").split(",")]:
    os.environ["TTN_MMA_DEEP"] = str(b)
    f = bench.build_workload(2)[0]
    t0 = time.perf_counter()
    plan = f.plan()
    tp = time.perf_counter() - t0
    best = 1e9
    for _ in range(6):
        o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
        best = min(best, o.kernel_ms)
    print(f"DEEP={b:2d}: plan {tp:6.2f} s  kernel {best:8.3f} ms  {npts / best / 1e6:7.3f} G pts/s  executed {o.flops_executed / npts:6.0f} flop/pt "
          f"= {o.flops_executed / best / 1e9:6.2f} TF", flush=True)
    del plan, f
