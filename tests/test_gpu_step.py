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
