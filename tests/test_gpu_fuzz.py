"""Randomised parity sweep (GPU): random chains and trees, random bond dimensions (uniform and
ragged), bases 2/3/4, real and complex index maps, one or two site indices per vertex — every
kernel that accepts the network against the 80-bit oracle, digits bit-exact."""
import numpy as np
import pytest

import cases
import itna_b200 as t
import oracle as orc
from itna_b200 import _capi

pytestmark = pytest.mark.gpu


def _random_network(seed):
    rng = np.random.default_rng(seed)
    kind = rng.choice(["mps", "comb_chain", "bintree", "unitree", "cmps"])
    base = int(rng.choice([2, 2, 2, 3, 4]))
    chi = int(rng.choice([1, 2, 3, 5, 8, 11, 16, 20, 32, 40]))
    if kind == "mps":
        L = int(rng.integers(2, 14))
        s = t.continuous_siteinds(t.named_grid((L, 1)), map_dimension=int(rng.integers(1, 4)) if L >= 3 else 1, base=base)
    elif kind == "comb_chain":
        ny = int(rng.integers(2, 8))
        g = t.named_comb_tree((2, ny))
        s = t.continuous_siteinds(g, [[(i, j) for j in range(1, ny + 1)] for i in (1, 2)], base=base)
    elif kind == "bintree":
        depth = int(rng.integers(2, 5))
        g = t.named_binary_tree(depth)
        vs = g.vertices()[int(rng.integers(0, 2)):]
        md = int(rng.integers(1, 4))
        s = t.continuous_siteinds(g, [vs[i::md] for i in range(md)], base=base)
        chi = min(chi, 20)
    elif kind == "unitree":
        n = int(rng.integers(3, 10))
        g = t.uniform_tree(n, rng=seed).rename_vertices(lambda v: (v, 1))
        s = t.continuous_siteinds(g, map_dimension=int(rng.integers(1, 3)), base=base)
        chi = min(chi, 5)
    else:
        L = int(rng.integers(2, 9))
        s = t.complex_continuous_siteinds(t.named_grid((L, 1)), map_dimension=int(rng.integers(1, 3)) if L >= 2 else 1,
                                          base=2 if base == 4 else base)
        chi = min(chi, 24)
    cplx_tensors = bool(rng.integers(0, 2)) or isinstance(s.indexmap, t.ComplexIndexMap)
    f = t.rand_itn(s, link_space=chi, rng=seed, eltype=complex if cplx_tensors else float, normalise=True)
    if rng.random() < 0.3 and chi <= 16:   # ragged link dimensions via a direct sum with a chi = 2 function
        f = f + (t.cos_itn(s, k=0.7, a=0.2, c=0.5) if cplx_tensors else t.cosh_itn(s, k=0.7, a=0.2, c=0.5))
    return kind, f


@pytest.mark.parametrize("seed", range(60))
def test_random_network_all_kernels(seed, monkeypatch):
    # chains: rotate the plan-time shape of the DMMA kernel's image over the seeds — default rule (short
    # chains collapse into two tables), uniform groups (every position a DMMA round), small deep tables
    if seed % 3 == 1:
        monkeypatch.setenv("TTN_MMA_DEEP", "0")
    elif seed % 3 == 2:
        monkeypatch.setenv("TTN_MMA_DEEP", "5")
    kind, f = _random_network(seed)
    imap = f.indexmap
    dims = imap.dimensions()
    rng = np.random.default_rng(1000 + seed)
    L = 8
    if isinstance(imap, t.ComplexIndexMap):
        pts = cases.complex_points(L, len(dims), rng, 120)
    else:
        pts = cases.edge_points(L, len(dims), rng, 160)
    plan = f.plan(dims)
    packed = plan.packed
    if packed.complex_coords:
        z = pts.astype(np.complex128)
        coords = np.empty((z.shape[0], 2 * z.shape[1]))
        coords[:, 0::2], coords[:, 1::2] = z.real, z.imag
    else:
        coords = pts.astype(np.float64)
    assert (plan.digits_host(coords) == orc.digits(packed, coords)).all()
    ref = orc.evaluate(packed, coords, orc.ORACLE_LD)
    avail = plan.info()["kernels_available"]
    ran = []
    for name, kid in _capi.KERNEL_IDS.items():
        if kid == 0 or name == "grid" or not (avail & (1 << kid)):  # "grid" only applies to ttn_evaluate_grid
            continue
        got, o = plan.evaluate_host(coords, kernel=name)
        err = orc.error_metric(got, ref).max()
        assert err < 1e-12, (seed, kind, name, err)
        ran.append(name)
    assert "generic" in ran
