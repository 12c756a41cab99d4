"""Compare the values the Julia reference returned (julia/ref_evaluate.jl) with libttneval on the same network
file and points:   python scripts/compare_julia_reference.py net.ttn.json values.json [--cpu-oracle]
Bars (north star): |v - v_ref| / max(|v_ref|, 1e-3 rms) <= 1e-12; on long random chains the reference's own FP64
arithmetic (BP + exp(sum log)) has a tail of a few 1e-12 against an 80-bit contraction (DESIGN.md 'Accuracy'), so
the exit status is p99.9 <= 1e-12 and max <= 1e-11, and both numbers are printed.  --cpu-oracle checks the file against the CPU
oracle instead of the GPU library (no GPU needed; test infrastructure only)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import itna_b200 as t

net, vals = sys.argv[1], sys.argv[2]
f, dims, pts = t.load_ttn(net)
ref = json.load(open(vals))
v_ref = np.array([complex(a, b) for a, b in ref["values"]])
pts = pts[: len(v_ref)]
if "--cpu-oracle" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    packed = t.pack(f, dims)
    if packed.complex_coords:
        c = np.empty((len(pts), 2 * pts.shape[1]))
        c[:, 0::2], c[:, 1::2] = pts.real, pts.imag
    else:
        c = pts
    got = orc.evaluate(packed, c, orc.ORACLE_LD)
else:
    got = t.evaluate(f, pts, dims)
if not np.iscomplexobj(got):
    print(f"max |imag| of the reference values (real network): {np.abs(v_ref.imag).max():.2e}")
    v_ref = v_ref.real
scale = np.sqrt(np.mean(np.abs(v_ref) ** 2))
err = np.abs(got - v_ref) / np.maximum(np.abs(v_ref), 1e-3 * scale)
print(f"{len(v_ref)} points: max floored rel err {err.max():.3e}, p99.9 {np.quantile(err, 0.999):.3e}, median {np.median(err):.3e}; "
      f"reference speed {ref.get('points_per_s', float('nan')):.1f} points/s on {ref.get('julia_threads', '?')} Julia thread(s)")
sys.exit(0 if (np.quantile(err, 0.999) <= 1e-12 and err.max() <= 1e-11) else 1)
