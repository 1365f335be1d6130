"""Receivers with more than 8 incoming edges against the LIVE reference (tests/golden/forward_longrows.npz, written by
oracle/gen_golden.py::gen_long_rows): cap 12 (the fused message kernel's loop beyond the first 8-edge chunk), an
uncapped C1-shaped batch at r = 5 (E/N ~ 88) and a 2 x 200-atom supercell batch at r = 7 uncapped (E/N ~ 148, the
"dense image neighbour lists" of BASELINE configs[2]; rows of up to 207 edges).  Per-layer x1 / x2 / h of selected
atoms (first, last, longest rows) and the model outputs; fp32 path at 1e-4, fp16 tensor path (fused kernel for the
capped case, CTA-per-atom gather + tensor-core fiber conv pair for the uncapped ones) at its stated tolerance."""
import numpy as np
import pytest
import torch

from conftest import TOL_FP16_MODEL, TOL_FP32, rel_err

pytestmark = pytest.mark.gpu
CASES = ["c1_cap12", "c1_uncapped", "c3_uncapped"]


def _engine(device, f, case, weights_npz, precision):
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    from arreau_b200.weights import PonitaWeights
    p = case + "/"
    sd = {k: weights_npz[k] for k in weights_npz.files if k not in ("ori_grid", "fourier_w")}
    w = PonitaWeights(sd, weights_npz["ori_grid"], device=device)
    eng = DenoiseEngine(w, build_tables(1000, 90), weights_npz["fourier_w"], f[p + "num_atoms"], float(f[p + "radius"]),
                        int(f[p + "cap"]), precision=precision, debug=True, device=device)
    eng.set_state(f[p + "frac"], f[p + "types"], f[p + "lengths"], f[p + "angles"])
    return eng


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_long_rows_against_reference(device, gold, weights_npz, case, precision):
    f = gold("forward_longrows.npz")
    p = case + "/"
    eng = _engine(device, f, case, weights_npz, precision)
    score, logits, len0 = eng.predict_scores(int(f[p + "timestep"]))
    torch.cuda.synchronize()
    E = eng.num_edges()
    # edge list identical to the reference's (receiver-major (i, j, cell) order)
    assert E == f[p + "src"].shape[0]
    assert np.array_equal(eng.src[:E].cpu().numpy(), f[p + "src"]) and np.array_equal(eng.dst[:E].cpu().numpy(), f[p + "dst"])
    deg = np.bincount(f[p + "dst"], minlength=eng.N)
    assert deg.max() > 8                       # the point of this fixture
    if precision == "fp16":                    # which message-pass implementation the step selects (csrc/step.cu)
        assert (eng.edge_capacity > 16 * eng.N) == (int(f[p + "cap"]) <= 0)
    sel = torch.as_tensor(f[p + "sel"], device=device)
    tol = TOL_FP32 if precision == "fp32" else TOL_FP16_MODEL
    for l in range(5):
        if precision == "fp32":                # the fp16 path keeps the message sums in shared memory / fp16 tiles
            assert rel_err(eng.x1_debug[l][sel].cpu().numpy(), f[p + f"x1_{l}"]) < tol, (l, "x1")
        assert rel_err(eng.x2_debug[l][sel].cpu().numpy(), f[p + f"x2_{l}"]) < tol, (l, "x2")
        assert rel_err(eng.h_debug[l + 1][sel].cpu().numpy(), f[p + f"h_{l}"]) < tol, (l, "h")
    assert rel_err(score.cpu().numpy(), f[p + "score"]) < tol
    assert rel_err(logits.cpu().numpy(), f[p + "logits"]) < tol
    assert rel_err(len0.cpu().numpy(), f[p + "len0"]) < tol
