"""BASELINE config 5: complex 2-D MPS, 40 vertices, Real+Imag index per vertex (phys dim 4), chi = 128."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import itna_b200 as t
from itna_b200 import _capi
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 20
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dfma, dmma = C.c_double(), C.c_double()
_capi.check(_capi.lib().ttn_measure_fp64_peak(0, C.byref(dfma), C.byref(dmma)))
s = t.complex_continuous_siteinds(t.named_grid((40, 1)), map_dimension=2)
t0 = time.time()
f = t.rand_itn(s, link_space=chi, rng=5, eltype=complex, normalise=True)
plan = f.plan()
info = plan.info()
print(f"plan built in {time.time() - t0:.1f} s; kernel {_capi.KERNEL_NAMES[info['auto_kernel']]}, flops/pt {info['flops_per_point']:.0f}")
x = torch.rand((npts, 4), dtype=torch.float64, device="cuda:0")
out = torch.empty(2 * npts, dtype=torch.float64, device="cuda:0")
best = 1e9
for _ in range(3):
    o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
    best = min(best, o.kernel_ms)
tf = info["flops_per_point"] * npts / (best * 1e-3) / 1e12
print(f"cfg5 chi={chi}: {npts:.2e} pts {best:.2f} ms {npts / best / 1e3:.2f} Mpts/s {tf:.2f} TFLOP/s algorithmic, {o.flops_executed / (best * 1e-3) / 1e12:.2f} TFLOP/s executed ({100 * o.flops_executed / (best * 1e-3) / 1e12 / dmma.value:.1f}% of DMMA peak {dmma.value:.1f}), launches {o.n_launches}")
# parity on a sample against the 80-bit oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle as orc
idx = np.arange(0, npts, max(1, npts // 300))[:300]
xs = x.cpu().numpy()[idx]
ref = orc.evaluate(plan.packed, xs, orc.ORACLE_LD, nthreads=orc.max_threads())
got = out.cpu().numpy().view(np.complex128)[idx]
print("audited 300 points: max floored rel err", orc.error_metric(got, ref).max())
