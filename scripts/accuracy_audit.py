"""Per-config accuracy audit (">= 10^6-point audited sample per config"; VERDICT r1 next #1b).

    python scripts/accuracy_audit.py CONFIG [NPOINTS] [--refined]   ->  profiles/r02_accuracy_cfgCONFIG.json

On the GPU box, all host threads: the GPU values (through the C ABI, planner's kernel) and the CPU restatements of
the reference's arithmetic (plain FP64 leaf-to-root, two-way BP + exp(sum log)) are each compared with the 80-bit
long-double contraction of the identical packed tensors on the SAME seeded points.  Reported: median / p99 / p99.9 /
max of the floored metric |v - ref| / max(|ref|, 1e-3 rms) and of the unfloored relative error; digits compared as
integers.  The oracle is the checker here, nothing else.  Config 3 (33 Mflop and 130 MB of slices per point on the
CPU) is audited on fewer points; the count is in the JSON."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import itna_b200 as t  # noqa: E402
import oracle as orc  # noqa: E402
from itna_b200 import _capi  # noqa: E402

DEFAULT_N = {1: 1 << 20, 2: 1 << 20, 3: 1 << 16, 4: 1 << 20, 5: 1 << 20, 6: 1 << 20, 7: 1 << 20}


def stats(e):
    return {"median": float(np.median(e)), "p99": float(np.quantile(e, 0.99)), "p99.9": float(np.quantile(e, 0.999)),
            "max": float(e.max()), "frac_above_1e-12": float((e > 1e-12).mean())}


def both(v, ref):
    un = np.abs(v - ref) / np.where(np.abs(ref) == 0, 1.0, np.abs(ref))
    return {"floored": stats(orc.error_metric(v, ref)), "unfloored": stats(un)}


def main():
    cfg = int(sys.argv[1])
    rest = [a for a in sys.argv[2:] if not a.startswith("--")]
    n = int(float(rest[0])) if rest else DEFAULT_N[cfg]
    f, ncol, _, desc = bench.build_workload(cfg)
    plan = f.plan()
    packed = plan.packed
    rng = np.random.default_rng(1000 + cfg)
    pts = rng.random((n, ncol))
    th = orc.max_threads()
    out = {"config": cfg, "workload": desc, "points": n, "host_threads": th,
           "metric": "|v - ref| / max(|ref|, 1e-3 rms(ref)) (floored) and |v - ref| / |ref| (unfloored); ref = 80-bit "
                     "long-double leaf-to-root contraction of the identical packed tensors (oracle ORACLE_LD)"}
    t0 = time.perf_counter()
    got, o = plan.evaluate_host(pts)
    out["gpu_kernel"] = _capi.KERNEL_NAMES[o.kernel_used]
    out["gpu_flops_executed_per_point"] = o.flops_executed / n
    dig_ok = True
    for lo in range(0, n, 1 << 18):
        sl = slice(lo, min(n, lo + (1 << 18)))
        dig_ok = dig_ok and bool((plan.digits_host(pts[sl]) == orc.digits(packed, pts[sl])).all())
    out["digits_identical"] = dig_ok
    t1 = time.perf_counter()
    ref = orc.evaluate(packed, pts, orc.ORACLE_LD, nthreads=th)
    out["seconds_ld_oracle"] = time.perf_counter() - t1
    out["gpu_fp64"] = both(got, ref)
    gr, o_r = plan.evaluate_host(pts, accuracy="refined")
    out["gpu_refined"] = both(gr, ref)
    out["gpu_refined"]["points_re_evaluated"] = int(o_r.n_refined)
    t1 = time.perf_counter()
    out["cpu_fp64_leaf_to_root"] = both(orc.evaluate(packed, pts, orc.ORACLE_F64, nthreads=th), ref)
    out["cpu_bp_exp_sum_log (reference-style)"] = both(orc.evaluate(packed, pts, orc.ORACLE_BP, nthreads=th), ref)
    out["seconds_cpu_restatements"] = time.perf_counter() - t1
    out["seconds_total"] = time.perf_counter() - t0
    path = os.path.join(ROOT, "gpurun_out", f"r02_accuracy_cfg{cfg}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
