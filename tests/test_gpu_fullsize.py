"""BASELINE.json configs[1] and configs[2] AT FULL SIZE against the oracle.  Crystals are
independent (no edge crosses a crystal, every pooling is per crystal), so a slice of the batch is a valid batch of its
own: the GPU runs ONE denoise step of the whole batch (1024 x 40 atoms, 5 A; 256 x 200 atoms, 7 A; cap 8) and the
oracle (oracle/restatement.py, fp64, pinned to the live reference by oracle/gen_golden.py) re-computes two slices of
it -- the first crystals and the last crystals, so block/tile boundaries at both ends of every buffer are covered --
from the same state and the same noise.  fp32 path: 1e-4; fp16 tensor path (the headline path): its stated tolerance.
An uncapped supercell batch (dense image neighbour lists; reduced to 16 crystals so the kernel slab fits comfortably)
is checked the same way."""
import numpy as np
import pytest
import torch

from conftest import TOL_FP16_MODEL, TOL_FP32, f64_default, rel_err

pytestmark = pytest.mark.gpu
Z, T = 90, 1000


def _oracle_slice(weights_npz, radius, cap, frac, types, lengths, angles, na, t, z_len, z_frac, u):
    from oracle import restatement as R
    with f64_default():
        T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
        sd = {k: T64(weights_npz[k]) for k in weights_npz.files if k not in ("ori_grid", "fourier_w")}
        W = R.PonitaWeights(sd, T64(weights_npz["ori_grid"]), radius)
        return R.denoise_step(W, R.DiffusionTables.build(T, Z), T64(weights_npz["fourier_w"]), T64(frac),
                              torch.as_tensor(types), T64(lengths), T64(angles), torch.as_tensor(na), t, T64(z_len),
                              T64(z_frac), T64(u), radius, cap)


def _wrapped(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b))
    return np.minimum(d, 1 - d).max()


def _check(device, packed_weights, weights_npz, G, n, radius, cap, t, precision, slice_crystals, seed):
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.synthetic import make_crystals
    from arreau_b200.tables import build_tables
    # teacher-forced state: Alexandria-shaped crystals noised to timestep t like the reference's forward process
    import bench
    frac, types, lengths, angles = bench.teacher_state(G, n, seed, t)
    eng = DenoiseEngine(packed_weights, build_tables(T, Z), weights_npz["fourier_w"], [n] * G, radius, cap,
                        precision=precision, device=device)
    eng.set_state(frac, types, lengths, angles)
    eng.draw_noise(seed, 0)
    z_len, z_frac, u = eng.z_len.cpu().numpy(), eng.z_frac.cpu().numpy(), eng.u_type.cpu().numpy()
    eng.step(t)
    torch.cuda.synchronize()
    assert int(eng.overflow_flag.item()) == 0
    E = eng.num_edges()
    row_ptr = eng.row_ptr.cpu().numpy()
    tol = TOL_FP32 if precision == "fp32" else TOL_FP16_MODEL
    got = dict(score=eng.score.cpu().numpy(), logits=eng.logits.cpu().numpy(), len0=eng.len0.cpu().numpy(),
               frac=eng.frac.cpu().numpy(), lengths=eng.lengths.cpu().numpy(), types=eng.types.cpu().numpy())
    worst = {}
    for g0 in (0, G - slice_crystals):
        cs, as_ = slice(g0, g0 + slice_crystals), slice(g0 * n, (g0 + slice_crystals) * n)
        ref = _oracle_slice(weights_npz, radius, cap, frac[as_], types[as_], lengths[cs], angles[cs], [n] * slice_crystals, t,
                            z_len[cs], z_frac[as_], u[as_])
        r_frac, r_types, r_len, _, r_score, r_logits, r_len0 = [x.numpy() for x in ref]
        # the GPU's edge count over this slice equals the oracle graph's
        from oracle import restatement as R
        with f64_default():
            lat = R.lattice_from_params(torch.as_tensor(lengths[cs]), torch.as_tensor(angles[cs]))
            ei = R.radius_graph_pbc(R.frac_to_cart_coords(torch.as_tensor(frac[as_]), lat, torch.full((slice_crystals,), n)),
                                    lat, torch.full((slice_crystals,), n), radius, cap)[0]
        assert int(row_ptr[as_.stop] - row_ptr[as_.start]) == ei.shape[1]
        src = eng.src[row_ptr[as_.start]:row_ptr[as_.stop]].cpu().numpy() - as_.start
        assert np.array_equal(src, ei[0].numpy())
        for name, r, sl in (("score", r_score, as_), ("logits", r_logits, as_), ("len0", r_len0, cs), ("lengths", r_len, cs)):
            e = rel_err(got[name][sl], r)
            worst[name] = max(worst.get(name, 0.0), e)
            assert e < tol, (precision, g0, name, e)
        assert _wrapped(got["frac"][as_], r_frac) < tol * max(1.0, np.abs(r_score).max())
        assert (got["types"][as_] != r_types).mean() <= 0.02
    print(f"full-size {G}x{n} r={radius} cap={cap} {precision}: E/N={E / (G * n):.2f} worst rel err {worst}")


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_c2_full_size_step_against_chunked_oracle(device, packed_weights, weights_npz, precision):
    _check(device, packed_weights, weights_npz, 1024, 40, 5.0, 8, 300, precision, slice_crystals=24, seed=11)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_c3_full_size_step_against_chunked_oracle(device, packed_weights, weights_npz, precision):
    _check(device, packed_weights, weights_npz, 256, 200, 7.0, 8, 300, precision, slice_crystals=2, seed=12)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_c3_uncapped_step_against_chunked_oracle(device, packed_weights, weights_npz, precision):
    """Dense image neighbour lists: 16 x 200 atoms, 7 A, uncapped (E/N ~ 75-150 -> the fp16 path's long-row message
    pass); oracle on the first and the last crystal."""
    _check(device, packed_weights, weights_npz, 16, 200, 7.0, 0, 300, precision, slice_crystals=1, seed=13)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_largest_dataset_cell_against_oracle(device, packed_weights, weights_npz, precision):
    """The dataset's largest cell (236 atoms, exploration/largest_system_in_dataset.py:34): 3 such crystals, 5 A, cap 8 --
    the image-culling graph instantiation, a tile count that is not a multiple of anything -- every crystal against
    the oracle."""
    _check(device, packed_weights, weights_npz, 3, 236, 5.0, 8, 400, precision, slice_crystals=1, seed=14)


def test_empty_batch_and_empty_graph(device, packed_weights, weights_npz):
    """Degenerate inputs: a batch without crystals is a no-op on every entry point; a batch whose atoms have no neighbour
    at all (one atom per 1000 A^3 cell, nothing within the cutoff, so E = 0) still runs the whole step and agrees with
    the oracle (the network then sees only the embedding)."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    tabs = build_tables(T, Z)
    for prec in ("fp32", "fp16"):
        eng = DenoiseEngine(packed_weights, tabs, weights_npz["fourier_w"], [], 5.0, 8, precision=prec, device=device)
        assert eng.N == 0 and eng.G == 0
        eng.set_state(np.zeros((0, 3)), np.zeros(0, dtype=np.int64), np.zeros((0, 3)), np.zeros((0, 3)))
        eng.draw_noise(1, 0)
        eng.step(500)
        torch.cuda.synchronize()
        assert eng.frac.shape == (0, 3) and eng.num_edges() == 0
    na = [1, 1, 1]
    rng = np.random.default_rng(2)
    frac, types = rng.random((3, 3)), rng.integers(0, 89, 3)
    lengths, angles = np.full((3, 3), 10.0) + rng.random((3, 3)), np.full((3, 3), np.pi / 2) + 0.05 * rng.standard_normal((3, 3))
    for prec, tol in (("fp32", TOL_FP32), ("fp16", TOL_FP16_MODEL)):
        eng = DenoiseEngine(packed_weights, tabs, weights_npz["fourier_w"], na, 5.0, 8, precision=prec, device=device)
        eng.set_state(frac, types, lengths, angles)
        eng.draw_noise(3, 0)
        z_len, z_frac, u = eng.z_len.cpu().numpy(), eng.z_frac.cpu().numpy(), eng.u_type.cpu().numpy()
        eng.step(500)
        torch.cuda.synchronize()
        assert eng.num_edges() == 0
        ref = _oracle_slice(weights_npz, 5.0, 8, frac, types, lengths, angles, na, 500, z_len, z_frac, u)
        assert rel_err(eng.score.cpu().numpy(), ref[4].numpy()) < tol
        assert rel_err(eng.logits.cpu().numpy(), ref[5].numpy()) < tol
        assert rel_err(eng.lengths.cpu().numpy(), ref[2].numpy()) < tol
        assert np.array_equal(eng.types.cpu().numpy(), ref[1].numpy())
