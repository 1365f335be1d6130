"""The training-step oracle (oracle/training.py) against the fixtures generated from the LIVE reference
(tests/golden/train_c5small.npz, written by oracle/gen_golden.py::gen_training).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import f64_default, rel_err


def _case(z, case):
    p = f"{case}/"
    return {k[len(p):]: z[k] for k in z.files if k.startswith(p)}


def _weights(weights_npz, radius=5.0):
    from oracle import restatement as R
    sd = {k: torch.as_tensor(weights_npz[k], dtype=torch.float64) for k in weights_npz.files
          if k not in ("ori_grid", "fourier_w")}
    return (R.PonitaWeights(sd, torch.as_tensor(weights_npz["ori_grid"], dtype=torch.float64), radius),
            torch.as_tensor(weights_npz["fourier_w"], dtype=torch.float64))


@pytest.mark.parametrize("case", [0, 1])
def test_training_loss_and_grads_match_reference(gold, weights_npz, case):
    from oracle import restatement as R, training as TR
    c = _case(gold("train_c5small.npz"), case)
    with f64_default():
        W, fw = _weights(weights_npz)
        tabs = R.DiffusionTables.build(1000, 90)
        T = lambda a: torch.as_tensor(a)  # noqa: E731
        loss, grads, parts = TR.training_grads(W, tabs, fw, T(c["frac0"]), T(c["types0"]), T(c["lattice0"]),
                                               T(c["num_atoms"]), T(c["timestep"]), T(c["eps_x"]), T(c["u"]),
                                               T(c["eps_l"]), 5.0, 8)
    assert abs(loss.item() - float(c["loss"])) <= 1e-11 * max(1.0, abs(float(c["loss"])))
    assert np.array_equal(parts["noisy_types"].numpy(), c["noisy_types"])
    for k in ("noisy_frac", "target_eps", "noisy_lengths", "pred_eps", "pred_logits", "pred_len"):
        assert rel_err(parts[k].numpy(), c[k]) < 1e-10, k
    if case == 0:
        for k, g in grads.items():
            assert rel_err(g.numpy(), c["grad/" + k]) < 1e-6, k          # fixtures stored as fp32
    else:
        for k, g in grads.items():
            assert abs(g.norm().item() - float(c["gradnorm/" + k])) <= 1e-9 * max(float(c["gradnorm/" + k]), 1e-30), k
            assert rel_err(g.reshape(-1)[:64].numpy(), c["gradhead/" + k]) < 1e-9 or float(c["gradnorm/" + k]) == 0.0, k


def test_noise_draw_order_is_the_references(gold):
    """timestep, eps_x, u, eps_l come out of ONE torch stream in the reference's order (diffusion_loss.py:214-237)."""
    from oracle import training as TR
    c = _case(gold("train_c5small.npz"), 0)
    with f64_default():
        torch.manual_seed(500)
        t, ex, u, el = TR.draw_training_noise(len(c["num_atoms"]), int(c["num_atoms"].sum()), 90, 1000)
    assert np.array_equal(t.numpy(), c["timestep"]) and np.array_equal(ex.numpy(), c["eps_x"])
    assert np.array_equal(u.numpy(), c["u"]) and np.array_equal(el.numpy(), c["eps_l"])


def test_calibrate_matches_reference(gold, weights_npz):
    from oracle import restatement as R, training as TR
    z = gold("train_c5small.npz")
    c = _case(z, "cal")
    with f64_default():
        W, fw = _weights(weights_npz)
        tabs = R.DiffusionTables.build(1000, 90)
        T = lambda a: torch.as_tensor(a)  # noqa: E731
        na, lengths, angles, frac = T(c["num_atoms"]), T(c["lengths"]), T(c["angles"]), T(c["frac"])
        G, N = na.shape[0], frac.shape[0]
        t = torch.full((N,), int(c["timestep"]), dtype=torch.long)
        lat = R.lattice_from_params(lengths, angles)
        rep = lambda a: torch.repeat_interleave(a, na, dim=0)  # noqa: E731
        x = torch.cat([torch.nn.functional.one_hot(T(c["types"]), 90), R.fourier_time_embedding(tabs.vp_betas[t].view(-1, 1), fw),
                       rep(na).unsqueeze(-1), rep(lengths), rep(angles), rep((lengths / na.unsqueeze(-1)).abs())], dim=1)
        vec = torch.cat([frac.unsqueeze(1), rep(lat)], dim=1)
        b = torch.repeat_interleave(torch.arange(G), na)
        ei, _, _, dist, direction = R.radius_graph_pbc(R.frac_to_cart_coords(frac, lat, na), lat, na, 5.0, 8)
        new = TR.calibrate(W, x, vec, ei, dist, direction, lat, b, G)
    n = 0
    for k in c:
        if k.startswith("after/"):
            assert rel_err(new[k[6:]].numpy(), c[k]) < 1e-6, k
            n += 1
    assert n == 10
