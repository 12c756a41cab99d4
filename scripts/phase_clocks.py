"""Per-phase clock accounting of the warp-autonomous DMMA kernel (debug build with
-DTTN_PHASE_CLOCKS, scripts/microbench/libttneval_dbg.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.pop("TTN_MMA_VARIANT", None)
import itna_b200 as t
from itna_b200 import _capi
_capi.LIB_PATH = os.path.join(ROOT, "scripts", "microbench", os.environ.get("TTN_DBG_LIB", "libttneval_dbg.so"))
import torch
L = _capi.lib()
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
f = t.rand_itn(s, link_space=16, rng=0, normalise=True)
plan = f.plan()
x = torch.rand((npts, 2), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
buf = (C.c_ulonglong * 8)()
plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="dmma")
L.ttn_debug_phase_clocks(buf, 1)
o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="dmma")
L.ttn_debug_phase_clocks(buf, 1)
names = ["digits+leaf", "sort (rank+list)", "ring wait", "batches (gather+DMMA+scatter)", "root", "class extract"]
tot = sum(buf[i] for i in range(6))
nw = buf[7]
print(f"kernel {o.kernel_ms:.3f} ms, warps {nw}, total warp-clk {tot:.3e}, per warp {tot / nw:.0f} clk")
subs = npts / 128
for i in range(6):
    print(f"{names[i]:32s} {100 * buf[i] / tot:5.1f}%   {buf[i] / subs:10.0f} clk per sub-tile   {buf[i] / subs / 29:8.0f} clk per round")
