"""Whole-step parity: one denoise step (one C call) against the live reference's goldens, teacher-forced at
t in {999, 750, 500, 250, 2, 1}, and the reference's real sample() (C1: 16 crystals, T=11) step by step."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-4


def _wrapped(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b))
    return np.minimum(d, 1 - d).max()


@pytest.mark.parametrize("timestep", [999, 750, 500, 250, 2, 1])
def test_teacher_forced_step_fp32(device, gold, packed_weights, weights_npz, timestep):
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    s = gold("steps_c1_T1000.npz")
    eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], s["num_atoms"], 5.0, 8,
                        device=device)
    p = f"t{timestep}/"
    eng.set_state(s[p + "frac"], s[p + "types"], s[p + "lengths"], s["angles"])
    # replay the reference's noise (the fixture stores u in fp32; regenerate the fp64 draw from the seed)
    si = [999, 750, 500, 250, 2, 1].index(timestep)
    torch.manual_seed(2000 + si)
    G, N = eng.G, eng.N
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        z_len, z_frac, u = torch.randn(G, 3), torch.randn(N, 3), torch.rand(N, 90)
    finally:
        torch.set_default_dtype(prev)
    assert np.array_equal(z_len.numpy(), s[p + "z_len"]) and np.array_equal(z_frac.numpy(), s[p + "z_frac"])
    eng.set_noise(z_len, z_frac, u)
    eng.step(timestep)
    torch.cuda.synchronize()
    assert rel_err(eng.score.cpu().numpy(), s[p + "score"]) < TOL_FP32
    assert rel_err(eng.logits.cpu().numpy(), s[p + "logits"]) < TOL_FP32
    assert rel_err(eng.len0.cpu().numpy(), s[p + "len0"]) < TOL_FP32
    assert rel_err(eng.lengths.cpu().numpy(), s[p + "lengths_next"]) < TOL_FP32
    assert rel_err(eng.lattice.cpu().numpy(), s[p + "lattice_next"]) < TOL_FP32
    # frac: |delta| = |score error| * (sigma_t^2 - sigma_{t-1}^2) <= 1e-4 * max|score| * 1
    assert _wrapped(eng.frac.cpu().numpy(), s[p + "frac_next"]) < TOL_FP32 * max(1.0, np.abs(s[p + "score"]).max())
    got, ref = eng.types.cpu().numpy(), s[p + "types_next"]
    # Gumbel-argmax on fp32 logits: identical unless the reference's own top-2 margin is within the logit error
    assert (got != ref).mean() <= 0.01


def test_reference_sample_T11_teacher_forced(device, gold, weights_npz):
    """C1: the reference's own sample() (16 crystals x 6 atoms, T = 11 -> 10 steps, calibrated length read-out);
    every step is checked from the reference's state entering that step."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.synthetic import calibrate_length_readout
    from arreau_b200.tables import build_tables
    from arreau_b200.weights import PonitaWeights
    s = gold("sample_T11.npz")
    n_per, G = int(s["n_per"]), int(s["num_crystals"])
    sd = {k: weights_npz[k].astype(np.float64) for k in weights_npz.files if k not in ("ori_grid", "fourier_w")}
    w = PonitaWeights(calibrate_length_readout(sd, n_per), weights_npz["ori_grid"], device=device)
    eng = DenoiseEngine(w, build_tables(11, 90), weights_npz["fourier_w"], [n_per] * G, 5.0, 8, device=device)
    steps = s["step_frac"].shape[0]
    for k, timestep in enumerate(reversed(range(1, 11))):
        eng.set_state(s["step_frac"][k], s["step_types"][k], s["step_lengths"][k], s["angles"])
        eng.set_noise(s["z_len"][k], s["z_frac"][k], s["u_type"][k])
        eng.step(timestep)
        torch.cuda.synchronize()
        assert rel_err(eng.score.cpu().numpy(), s["step_score"][k]) < TOL_FP32, timestep
        assert rel_err(eng.logits.cpu().numpy(), s["step_logits"][k]) < TOL_FP32, timestep
        assert rel_err(eng.len0.cpu().numpy(), s["step_len0"][k]) < TOL_FP32, timestep
        if k + 1 < steps:
            assert _wrapped(eng.frac.cpu().numpy(), s["step_frac"][k + 1]) < TOL_FP32 * max(1.0, np.abs(s["step_score"][k]).max())
            assert rel_err(eng.lengths.cpu().numpy(), s["step_lengths"][k + 1]) < TOL_FP32
            assert (eng.types.cpu().numpy() != s["step_types"][k + 1]).mean() <= 0.02
        else:
            assert _wrapped(eng.frac.cpu().numpy(), s["frac_x"]) < TOL_FP32 * max(1.0, np.abs(s["step_score"][k]).max())
            assert rel_err(eng.lattice.cpu().numpy(), s["lattice"]) < TOL_FP32


def test_staged_noise_equals_direct(device, gold, packed_weights, weights_npz):
    """engine.stage_noise / use_staged_noise (next step's draws uploaded on a side stream into a second buffer set)
    gives bit-identical steps to set_noise."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.tables import build_tables
    s = gold("steps_c1_T1000.npz")
    g = torch.Generator().manual_seed(11)

    def run(staged):
        eng = DenoiseEngine(packed_weights, build_tables(1000, 90), weights_npz["fourier_w"], s["num_atoms"], 5.0, 8,
                            device=device)
        eng.set_state(s["t500/frac"], s["t500/types"], s["t500/lengths"], s["angles"])
        g.manual_seed(11)
        noises = [(torch.randn(eng.G, 3, generator=g, dtype=torch.float64).pin_memory(),
                   torch.randn(eng.N, 3, generator=g, dtype=torch.float64).pin_memory(),
                   torch.rand(eng.N, 90, generator=g, dtype=torch.float64).pin_memory()) for _ in range(4)]
        if staged:
            eng.stage_noise(*noises[0])
        for k, t in enumerate((500, 499, 498, 497)):
            if staged:
                eng.use_staged_noise()
            else:
                eng.set_noise(*noises[k])
            eng.step(t)
            if staged:
                eng.release_noise()
                if k + 1 < 4:
                    eng.stage_noise(*noises[k + 1])
        torch.cuda.synchronize()
        return eng.frac.clone(), eng.types.clone(), eng.lengths.clone()

    a, b = run(False), run(True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_sample_draws_the_reference_noise_stream(device, gold, weights_npz):
    """DiffusionLoss.sample in its default host-noise mode draws numpy / torch CPU random numbers exactly as the
    reference's sample() does (diffusion_loss.py:294-316, then per step randn_like(lengths), randn_like(frac),
    rand((N, Z)): SURVEY 3.1): with the seeds of tests/golden/sample_T11.npz the initial angles and every step's three
    noise tensors are BIT-IDENTICAL to what the reference consumed.  The free-running trajectory itself is chaotic in
    the sampler's ~1 A initial cells (exactly tied periodic images flip on 1-ulp differences, see
    oracle/gen_golden.py), so the end state is compared loosely; step-by-step parity is the teacher-forced test above."""
    import argparse
    from arreau_b200.diffusion.diffusion_loss import DiffusionLoss
    from arreau_b200.ponita.models.ponita import PonitaFiberBundle
    from arreau_b200.synthetic import calibrate_length_readout
    s = gold("sample_T11.npz")
    n_per, G = int(s["n_per"]), int(s["num_crystals"])
    sd = calibrate_length_readout({k: weights_npz[k].astype(np.float64) for k in weights_npz.files
                                   if k not in ("ori_grid", "fourier_w")}, n_per)
    net = PonitaFiberBundle((164, 4), 128, 90, 3, 0, 0, 5, output_dim_vec=1, radius=5.0, num_ori=16, basis_dim=256, degree=3,
                            widening_factor=4, layer_scale=1e-6, multiple_readouts=True, ori_grid=weights_npz["ori_grid"])
    net.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    dl = DiffusionLoss(argparse.Namespace(radius=5.0, max_neighbors=8, num_timesteps=11), 90)
    seen = []

    def spy(timestep, eng):
        seen.append((timestep, eng.z_len.cpu().numpy().copy(), eng.z_frac.cpu().numpy().copy(),
                     eng.u_type.cpu().numpy().copy(), eng.frac.cpu().numpy().copy(), eng.lengths.cpu().numpy().copy()))

    np.random.seed(77)
    torch.manual_seed(77)
    zt = list(range(1, 90)) + [2001]
    res = dl.sample(model=net, z_table=zt, t_emb_weights=torch.as_tensor(weights_npz["fourier_w"]),
                    num_atoms_per_sample=n_per, num_samples_in_batch=G, device=device, step_callback=spy)
    assert [k[0] for k in seen] == list(reversed(range(1, 11)))
    eng = dl._engine
    assert np.array_equal(eng.angles.cpu().numpy(), s["angles"])            # numpy stream: sample_bravais_angles
    for k, (_, zl, zf, uu, _, _) in enumerate(seen):
        assert np.array_equal(zl, s["z_len"][k]), k                          # torch CPU stream, reference order
        assert np.array_equal(zf, s["z_frac"][k]), k
        assert np.array_equal(uu, s["u_type"][k]), k
    # first step: same initial state (lengths0 / frac0 are the draws before the loop) -> step 1 output within fp32 error
    assert _wrapped(seen[0][4], s["step_frac"][1]) < TOL_FP32 * max(1.0, np.abs(s["step_score"][0]).max())
    assert rel_err(seen[0][5], s["step_lengths"][1]) < TOL_FP32
    # end of the free-running trajectory: same crystals up to the chaos of the degenerate early cells
    assert res.num_atoms.tolist() == s["num_atoms"].tolist() and res.frac_x.shape == s["frac_x"].shape
    lat_err = rel_err(res.lattice, s["lattice"])
    frac_dev = np.abs(res.frac_x - s["frac_x"]); frac_dev = np.minimum(frac_dev, 1 - frac_dev)
    same_types = float((res.atomic_numbers == s["atomic_numbers"]).mean())
    print(f"free-running sample vs reference: lattice rel {lat_err:.2e}, frac max {frac_dev.max():.2e} "
          f"median {np.median(frac_dev):.2e}, types equal {same_types:.3f}")
    # measured on B200: lattice 2.8e-5, frac max 5.7e-5 / median 7e-9, every type equal
    assert lat_err < 1e-3 and np.median(frac_dev) < 1e-5 and same_types >= 0.97


def test_fp16_trajectory_tracks_fp32_trajectory(device, weights_npz):
    """A WHOLE 999-step trajectory on the fp16 tensor path against the fp32 path (same Philox noise, same initial
    state, 64 crystals x 40 atoms, calibrated length read-out): per-step rounding differences (~1e-3) and the
    occasional Gumbel-argmax flip compound over 999 steps, so individual atoms decorrelate; what a sampler user
    relies on are the trajectory's statistics: E/N trace, final cell lengths, final type histogram."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.synthetic import calibrate_length_readout
    from arreau_b200.tables import build_tables
    from arreau_b200.weights import PonitaWeights
    G, n = 64, 40
    sd = calibrate_length_readout({k: weights_npz[k] for k in weights_npz.files if k not in ("ori_grid", "fourier_w")}, n)
    w = PonitaWeights(sd, weights_npz["ori_grid"], device=device)
    rng = np.random.default_rng(3)
    angles = np.stack([np.full(G, 90.0), rng.uniform(90, 180, G), np.full(G, 90.0)], 1)
    lengths, frac = rng.standard_normal((G, 3)), rng.standard_normal((G * n, 3))
    out = {}
    for prec in ("fp32", "fp16"):
        eng = DenoiseEngine(w, build_tables(1000, 90), weights_npz["fourier_w"], [n] * G, 5.0, 8, precision=prec, device=device)
        eng.set_state(frac, np.full(G * n, 89), lengths, angles)
        epa = []
        for k, t in enumerate(reversed(range(1, 1000))):
            eng.draw_noise(5, k)
            eng.step(t)
            if k % 10 == 0:
                epa.append(eng.num_edges() / (G * n))
        torch.cuda.synchronize()
        out[prec] = dict(epa=np.asarray(epa), lengths=eng.lengths.cpu().numpy(), types=eng.types.cpu().numpy(),
                         frac=eng.frac.cpu().numpy())
        assert np.isfinite(out[prec]["lengths"]).all() and np.isfinite(out[prec]["frac"]).all()
    a, b = out["fp32"], out["fp16"]
    epa_dev = float(np.abs(a["epa"] - b["epa"]).max())
    len_rel = float(np.abs(a["lengths"] - b["lengths"]).max() / np.abs(a["lengths"]).max())
    ha, hb = np.bincount(a["types"], minlength=90) / a["types"].size, np.bincount(b["types"], minlength=90) / b["types"].size
    tv = 0.5 * float(np.abs(ha - hb).sum())
    same = float((a["types"] == b["types"]).mean())
    print(f"fp16 vs fp32 trajectory: max |E/N diff| {epa_dev:.3f} (E/N {a['epa'].min():.2f}..{a['epa'].max():.2f}), "
          f"final lengths rel diff {len_rel:.2e}, type histogram TV {tv:.3f}, identical types {same:.3f}")
    assert a["epa"].min() > 1.0                       # the graph never empties (calibrated read-out)
    # measured on B200: E/N identical at every sampled step, final lengths 1.5e-4 apart, identical final types
    assert epa_dev < 0.05 and len_rel < 5e-3 and tv < 0.02 and same > 0.95


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_cuda_graph_replay_is_bit_identical(device, packed_weights, weights_npz, precision):
    """arreau_denoise_step_replay captured once as a CUDA graph and replayed (device-resident step counter, timestep and
    VP coefficients; Philox noise inside the graph) against the ordinary loop `draw_noise(seed, k); step(t)`: the same
    bits after 40 steps, including the clamp at t = 1, and again after rewinding the counter on a new state."""
    from arreau_b200.engine import DenoiseEngine
    from arreau_b200.synthetic import make_crystals
    from arreau_b200.tables import build_tables
    cr = make_crystals(5, 3, 17, seed=4)
    tabs = build_tables(1000, 90)

    def run(graph, t_first, steps, seed, state):
        eng = DenoiseEngine(packed_weights, tabs, weights_npz["fourier_w"], cr.num_atoms, 5.0, 8, precision=precision, device=device)
        eng.set_state(*state)
        if graph:
            g = eng.capture_trajectory_graph(t_first, seed)
            assert int(eng._replay_bufs["counter"].item()) == 0
            for _ in range(steps):
                g.replay()
        else:
            for k in range(steps):
                eng.draw_noise(seed, k)
                eng.step(max(t_first - k, 1))
        torch.cuda.synchronize()
        return eng, [t.clone() for t in (eng.frac, eng.types, eng.lengths, eng.lattice, eng.score, eng.logits)]

    state = (cr.frac, cr.types, cr.lengths, cr.angles)
    for t_first, steps in ((700, 40), (20, 25)):           # the second run walks through t = 1 and stays there
        _, a = run(False, t_first, steps, 9, state)
        eng, b = run(True, t_first, steps, 9, state)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        assert int(eng._replay_bufs["counter"].item()) == steps
    # rewind: a second trajectory on the same engine and graph
    state2 = (cr.frac[::-1].copy(), cr.types, cr.lengths * 1.1, cr.angles)
    _, a = run(False, 500, 10, 9, state2)
    eng.set_state(*state2)
    eng.reset_trajectory_graph()
    eng._replay.t_first = 500          # the struct is read by value at capture: a new t_first needs a new capture
    g = eng.capture_trajectory_graph(500, 9)
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    for x, y in zip(a, (eng.frac, eng.types, eng.lengths, eng.lattice, eng.score, eng.logits)):
        assert torch.equal(x, y)
