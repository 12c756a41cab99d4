"""BASELINE config 3: 3-D binary tree depth 7 (127 vertices, 120 with a binary site index, dims interleaved),
chi = 64, real N(0,1)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import itna_b200 as t
from itna_b200 import _capi
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 17
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dfma, dmma = C.c_double(), C.c_double()
_capi.check(_capi.lib().ttn_measure_fp64_peak(0, C.byref(dfma), C.byref(dmma)))
g = t.named_binary_tree(7)
vs = g.vertices()
with_site = vs[7:]            # 120 vertices carry a site index, the top 7 do not
s = t.continuous_siteinds(g, [with_site[i::3] for i in range(3)])
t0 = time.time()
f = t.rand_itn(s, link_space=chi, rng=3, normalise=True)
plan = f.plan()
info = plan.info()
print(f"plan built in {time.time() - t0:.1f} s; kernel {_capi.KERNEL_NAMES[info['auto_kernel']]}, flops/pt {info['flops_per_point']:.0f}, tensors {info['tensor_bytes'] / 1e6:.0f} MB")
x = torch.rand((npts, 3), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
best = 1e9
for _ in range(2):
    o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
    best = min(best, o.kernel_ms)
tf = info["flops_per_point"] * npts / (best * 1e-3) / 1e12
print(f"cfg3 chi={chi}: {npts:.2e} pts {best:.2f} ms {npts / best / 1e3:.3f} Mpts/s {tf:.2f} TFLOP/s algorithmic, {o.flops_executed / (best * 1e-3) / 1e12:.2f} TFLOP/s executed ({100 * o.flops_executed / (best * 1e-3) / 1e12 / dmma.value:.1f}% of DMMA peak {dmma.value:.1f}; {o.flops_executed / npts:.0f} flop/pt), launches {o.n_launches}")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle as orc
idx = np.arange(0, npts, max(1, npts // 64))[:64]
xs = x.cpu().numpy()[idx]
ref = orc.evaluate(plan.packed, xs, orc.ORACLE_LD, nthreads=orc.max_threads())
got = out.cpu().numpy()[idx]
print("audited 64 points: max floored rel err", orc.error_metric(got, ref).max())
