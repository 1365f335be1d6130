"""K2..K7 parity: the CUDA Ponita forward against goldens computed by the LIVE reference in fp64
(tests/golden/forward_c1_t500.npz, written by oracle/gen_golden.py).  fp32 path: 1e-4 of max|ref|
(BASELINE.json north_star); the per-layer intermediates make the test sensitive to the conv / MLP
branches (quirk B6)."""
import types as _types

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-4


def _graph(device, f):
    g = _types.SimpleNamespace()
    g.x = torch.as_tensor(f["x"], device=device)
    g.vec = torch.as_tensor(f["vec"], device=device)
    g.edge_index = torch.as_tensor(f["edge_index"], device=device)
    g.dists = torch.as_tensor(f["dist"], device=device)
    g.inter_atom_direction = torch.as_tensor(f["direction"], device=device)
    g.lattice = torch.as_tensor(f["lattice"], device=device)
    g.batch = torch.as_tensor(f["batch"], device=device)
    return g


def _model(device, w, precision="fp32"):
    from arreau_b200.ponita.models.ponita import PonitaFiberBundle
    Z = 90
    m = PonitaFiberBundle((164, 4), 128, Z, 3, 0, 0, 5, output_dim_vec=1, radius=5.0, num_ori=16, basis_dim=256,
                          degree=3, widening_factor=4, layer_scale=1e-6, multiple_readouts=True,
                          ori_grid=w["ori_grid"], precision=precision)
    sd = {k: torch.as_tensor(w[k]) for k in w.files if k not in ("ori_grid", "fourier_w")}
    missing = m.load_state_dict(sd)
    assert all(k.endswith("callibrated") for k in missing.missing_keys), missing.missing_keys
    return m.to(device)


def test_forward_matches_reference_fp32(device, gold, weights_npz):
    f = gold("forward_c1_t500.npz")
    m = _model(device, weights_npz)
    logits, vec, len0, gv, edge = m(_graph(device, f))
    assert gv is None and edge == [None] * 5 and vec.shape[1:] == (1, 3)
    assert rel_err(logits.cpu().numpy(), f["logits"]) < TOL_FP32
    assert rel_err(vec.cpu().numpy(), f["vec_out"]) < TOL_FP32
    assert rel_err(len0.cpu().numpy(), f["len0"]) < TOL_FP32


def test_forward_accepts_unsorted_edges(device, gold, weights_npz):
    f = gold("forward_c1_t500.npz")
    m = _model(device, weights_npz)
    g = _graph(device, f)
    perm = torch.randperm(g.edge_index.shape[1], generator=torch.Generator().manual_seed(0)).to(device)
    g.edge_index, g.dists, g.inter_atom_direction = g.edge_index[:, perm], g.dists[perm], g.inter_atom_direction[perm]
    logits, vec, len0, _, _ = m(g)
    assert rel_err(logits.cpu().numpy(), f["logits"]) < TOL_FP32
    assert rel_err(len0.cpu().numpy(), f["len0"]) < TOL_FP32


def _engine_at_t500(device, gold, packed_weights, weights_npz, precision="fp32", debug=True):
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    s = gold("steps_c1_T1000.npz")
    eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], s["num_atoms"], 5.0, 8,
                        precision=precision, debug=debug, device=device)
    eng.set_state(s["t500/frac"], s["t500/types"], s["t500/lengths"], s["angles"])
    return eng, s


def test_layer_intermediates_fp32(device, gold, packed_weights, weights_npz):
    f = gold("forward_c1_t500.npz")
    eng, s = _engine_at_t500(device, gold, packed_weights, weights_npz)
    score, logits, len0 = eng.predict_scores(500)
    torch.cuda.synchronize()
    # the engine rebuilt the same inputs and graph from the raw state
    assert np.array_equal(eng.src[: eng.num_edges()].cpu().numpy(), f["edge_index"][0])
    assert rel_err(eng.x.cpu().numpy(), f["x"]) < 1e-6
    assert rel_err(eng.vec.cpu().numpy(), f["vec"]) < 1e-6
    n = f["h0_first"].shape[0]
    assert rel_err(eng.h_debug[0, :n].cpu().numpy(), f["h0_first"]) < 1e-5
    for l in range(5):
        assert rel_err(eng.x1_debug[l, :n].cpu().numpy(), f[f"x1_{l}_first"]) < TOL_FP32, l
        assert rel_err(eng.x2_debug[l, :n].cpu().numpy(), f[f"x2_{l}_first"]) < TOL_FP32, l
        assert rel_err(eng.h_debug[l + 1, :n].cpu().numpy(), f[f"h_{l}_first"]) < TOL_FP32, l
    assert rel_err(logits.cpu().numpy(), f["logits"]) < TOL_FP32
    assert rel_err(score.cpu().numpy(), f["vec_out"][:, 0]) < TOL_FP32
    assert rel_err(len0.cpu().numpy(), f["len0"]) < TOL_FP32


def test_kernel_basis_first_edges_fp32(device, gold, packed_weights, weights_npz):
    """kernels[l] = kernel_basis @ Wk_l^T for the first edges, against the reference's kernel_basis."""
    f = gold("forward_c1_t500.npz")
    eng, _ = _engine_at_t500(device, gold, packed_weights, weights_npz)
    eng.predict_scores(500)
    torch.cuda.synchronize()
    kb = f["kernel_basis_first_edges"].astype(np.float64)           # [8,16,256]
    for l in (0, 4):
        wk = weights_npz[f"interaction_layers.{l}.conv.kernel.weight"].astype(np.float64)
        ref = kb @ wk.T
        got = eng.kernels[l, : kb.shape[0]].float().cpu().numpy()
        assert rel_err(got, ref) < TOL_FP32, l


def test_forward_deterministic(device, gold, packed_weights, weights_npz):
    eng, _ = _engine_at_t500(device, gold, packed_weights, weights_npz, debug=False)
    a = [t.clone() for t in eng.predict_scores(500)]
    b = [t.clone() for t in eng.predict_scores(500)]
    for x, y in zip(a, b):
        assert torch.equal(x, y)
