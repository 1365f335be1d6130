"""PONITA_DIFFUSION mirror end to end on the GPU: sample(), generate_n_crystals + writer, and a short
training loop through training_step / configure_optimizers."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
Z = 90


def _args(T):
    return argparse.Namespace(dataset="synthetic", lr=1e-3, weight_decay=0.0, epochs=4, warmup=1, layer_scale=1e-6,
                              train_augm=False, hidden_dim=128, layers=5, radius=5.0, num_ori=16, basis_dim=256, degree=3,
                              widening_factor=4, multiple_readouts=True, num_timesteps=T, max_neighbors=8)


def _model(device, weights_npz, T, precision="fp32"):
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
    from arreau_b200.synthetic import calibrate_length_readout
    from arreau_b200.tools.atomic_number_table import AtomicNumberTable
    m = PONITA_DIFFUSION(_args(T), AtomicNumberTable(list(range(1, Z)) + [2001]), ori_grid=weights_npz["ori_grid"],
                         precision=precision)
    sd = calibrate_length_readout({k: weights_npz[k] for k in weights_npz.files if k not in ("ori_grid", "fourier_w")}, 6)
    m.model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    with torch.no_grad():
        m.t_emb.gaussian_fourier_proj_w.copy_(torch.as_tensor(weights_npz["fourier_w"]))
    return m.to(device)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_wrapper_sample_and_generate(device, weights_npz, tmp_path, precision):
    from arreau_b200.generate import generate_n_crystals
    from arreau_b200.inference.process_generated_crystals import load_sample_results_from_hdf5
    m = _model(device, weights_npz, T=21, precision=precision)
    np.random.seed(0)
    torch.manual_seed(0)
    res = m.sample(num_atoms_per_sample=6, num_samples_in_batch=8, device=device)
    assert res.frac_x.shape == (48, 3) and res.lattice.shape == (8, 3, 3) and res.atomic_numbers.shape == (48,)
    assert np.isfinite(res.frac_x).all() and (res.frac_x >= 0).all() and (res.frac_x < 1).all()
    assert np.isfinite(res.lattice).all() and set(np.unique(res.atomic_numbers)) <= set(list(range(1, Z)) + [2001])
    out = str(tmp_path / "crystals.h5")
    full = generate_n_crystals(m, 20, 6, None, num_crystals_per_batch=8, device=device, out_path=out)
    assert full.frac_x.shape == (120, 3) and full.num_atoms.tolist() == [6] * 20
    assert full.idx_start.tolist() == list(range(0, 120, 6))
    back = load_sample_results_from_hdf5(out if not out.endswith(".h5") else out)
    assert np.array_equal(back.frac_x, full.frac_x) and np.array_equal(back.lattice, full.lattice)
    # constant atoms (use_constant_atomic_symbols, diffusion_loss.py:348-349): types are not resampled
    res2 = m.sample(num_atoms_per_sample=2, num_samples_in_batch=3, use_constant_atomic_symbols=["C", "O"], device=device)
    assert res2.atomic_numbers.tolist() == [6, 6, 6, 8, 8, 8]


def test_wrapper_training_loop(device, weights_npz):
    """training_step -> loss.backward() -> fused Adam, with the cosine-warmup schedule: loss falls on a fixed batch."""
    from arreau_b200.diffusion.lattice_helpers import lattice_from_params
    from arreau_b200.synthetic import make_training_batch
    m = _model(device, weights_npz, T=1000)
    opt = m.configure_optimizers(device)
    cr = make_training_batch(24, seed=3)
    L0 = lattice_from_params(torch.as_tensor(cr.lengths).to(device), torch.as_tensor(cr.angles).to(device))
    batch = argparse.Namespace(X0=torch.as_tensor(cr.frac).to(device), A0=torch.as_tensor(cr.types).to(device),
                               L0=L0.reshape(-1, 3), num_atoms=torch.as_tensor(cr.num_atoms).to(device))
    losses = []
    for epoch in range(4):
        opt.set_epoch(epoch + 1, m.warmup, 8)
        for _ in range(3):
            torch.manual_seed(5)                       # same draws every step: a fixed noised batch
            loss = m.training_step(batch)
            for p in m.model.parameters():
                p.grad = None
            loss.backward()
            # .backward() hands the parameters the slices of the flat gradient buffer the step's kernels wrote
            assert torch.equal(m.model.x_embedder.weight.grad, m.model.flat.grad_views()["x_embedder.weight"])
            opt.step()
            losses.append(loss.item())
    assert losses[-1] < losses[0], losses


def test_fit_loop_on_a_tiny_dataset(device, weights_npz, tmp_path):
    """arreau_b200.train.fit (the training entry point without Lightning): dataset file -> collate -> callibrate ->
    steps with the TF32 backward -> epoch metric; the loss falls over a few epochs on a tiny dataset."""
    from arreau_b200.diffusion.lattice_dataset import CrystalDataset, save_dataset_npz
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
    from arreau_b200.synthetic import make_crystals
    from arreau_b200.train import default_args, fit
    cr = make_crystals(24, 2, 10, seed=9)
    off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
    zs = [cr.types[off[i]:off[i + 1]] % 5 + 1 for i in range(24)]
    frac = [cr.frac[off[i]:off[i + 1]] for i in range(24)]
    lat = np.stack([np.diag(cr.lengths[i]) for i in range(24)])
    ds = CrystalDataset([save_dataset_npz(str(tmp_path / "tiny"), zs, lat, frac)])
    torch.manual_seed(0)
    args = default_args(lr=2e-3, epochs=6, warmup=1, layer_scale=1e-6)
    model = PONITA_DIFFUSION(args, ds.z_table)
    hist = fit(model, ds, epochs=6, batch_size=12, device=device, backward_precision="tf32", log=lambda *_: None)
    assert len(hist) == 6 and all(np.isfinite(hist)) and min(hist[3:]) < hist[0], hist
    assert all(bool(l.conv.callibrated) for l in model.model.interaction_layers)
