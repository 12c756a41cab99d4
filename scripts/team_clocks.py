"""Per-phase clock accounting of the team-sorted DMMA kernel on config 2 (debug build with
-DTTN_TEAM_CLOCKS: `make -C itensornumericalanalysis.jl_b200/csrc dbgteam`)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itna_b200 as t
from itna_b200 import _capi
_capi.LIB_PATH = os.path.join(ROOT, "scripts", "microbench", "libttneval_teamdbg.so")
import torch
L = _capi.lib()
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000
base = int(sys.argv[2]) if len(sys.argv) > 2 else 2   # base 3 / 4: a 60-site MPS of one coordinate
if base == 2:
    g = t.named_comb_tree((2, 30))
    s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
else:
    s = t.continuous_siteinds(t.named_grid((60, 1)), base=base)
f = t.rand_itn(s, link_space=16, rng=0, normalise=True)
plan = f.plan()
x = torch.rand((npts, 2 if base == 2 else 1), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
buf = (C.c_ulonglong * 12)()
plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="dmma")
L.ttn_debug_team_clocks(buf, 1)
o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="dmma")
L.ttn_debug_team_clocks(buf, 1)
names = ["K1 (+ root prefetch)", "leaf issue + count round 0", "leaf wait (cp.async)", "barrier 0", "rounds: list + count next",
         "rounds: barrier", "rounds: classes (gather+DMMA+scatter)", "last barrier", "root + store"]
tot = sum(buf[i] for i in range(9))
nw = buf[11]
print(f"kernel {o.kernel_ms:.3f} ms for {npts} points, warps {nw}, per warp {tot / nw:.0f} clk")
tiles = npts / 128
for i in range(9):
    print(f"{names[i]:40s} {100 * buf[i] / tot:5.1f}%   {buf[i] / tiles:10.0f} clk per warp-tile")
