"""Table-kernel probe (GPU): launch variants x table budgets on the small-chi datapoints, next to the
register chain kernel the planner used for these networks before.  Prints kernel ms (CUDA events, best of 5),
points/s and algorithmic GB/s (8 B per coordinate + 8 / 16 B per value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t

npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
SWEEP = [("0", "200", "1"), ("0", "200", "0"), ("0", "100", "1")]   # (unused, table KB, replicated layout on / off)


def run(name, f, ncol):
    x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
    ref = None
    for variant, kb, rep in SWEEP + [("chain", "", ""), ("dmma", "", "")]:
        kernel = "table"
        if variant in ("chain", "dmma"):
            kernel = variant
        else:
            os.environ["TTN_TABLE_VARIANT"], os.environ["TTN_TABLE_KB"], os.environ["TTN_TABLE_REP"] = variant, kb, rep
        f._plans.clear()
        plan = f.plan()
        info = plan.info()
        n = npts if kernel == "table" else npts // 10
        out = torch.empty(n * (2 if info["is_complex"] else 1), dtype=torch.float64, device="cuda:0")
        best = 1e9
        for it in range(5):
            o = plan.evaluate_device(x.data_ptr(), n, out.data_ptr(), kernel=kernel)
            best = min(best, o.kernel_ms)
        if ref is None:
            ref = out.clone()
        dev = (out - ref[: out.numel()]).abs().max().item() / ref.abs().max().item()
        gbs = info["bytes_per_point"] * n / (best * 1e-3) / 1e9
        print(f"{name:22s} kernel={kernel:5s} variant={variant:5s} kb={kb:3s} rep={rep:1s} {n:.1e} pts {best:8.3f} ms "
              f"{n / best / 1e6:8.2f} G pts/s {gbs:8.1f} GB/s  (max dev from first {dev:.1e})", flush=True)
    os.environ.pop("TTN_TABLE_VARIANT", None)
    os.environ.pop("TTN_TABLE_KB", None)
    os.environ.pop("TTN_TABLE_REP", None)
    f._plans.clear()


g = t.named_comb_tree((2, 30))
s2 = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
run("exp chi1 2x30", t.exp_itn(s2, k=0.9, a=0.1, c=1.2, dim=1), 2)
run("rand chi2 2x30", t.rand_itn(s2, link_space=2, rng=20267, normalise=True), 2)
run("rand chi4 2x30", t.rand_itn(s2, link_space=4, rng=20268, normalise=True), 2)
s24 = t.continuous_siteinds(t.named_grid((24, 1)), map_dimension=3)
run("rand chi4 mps24 3-D", t.rand_itn(s24, link_space=4, rng=20271, normalise=True), 3)
sc = t.complex_continuous_siteinds(t.named_grid((30, 1)), map_dimension=1)
run("complex chi1 1-D 2x30", t.rand_itn(sc, link_space=1, rng=20269, eltype=complex, normalise=True), 2)
si = t.continuous_siteinds(t.named_grid((60, 1)), map_dimension=2)
run("exp chi1 mps60 interleaved", t.exp_itn(si, k=0.9, a=0.1, c=1.2, dim=2), 2)
run("rand chi2 mps60 interleaved", t.rand_itn(si, link_space=2, rng=20270, normalise=True), 2)
s1 = t.continuous_siteinds(t.named_grid((20, 1)))
run("sin_qtt20 (cfg1 net)", t.sin_itn(s1, k=2.0, a=0.3, c=1.1), 1)
