"""Checkpoint ingestion with a checkpoint PICKLED BY THE LIVE REFERENCE CLASSES
(tests/golden/reference_model.ckpt, written by oracle/gen_golden.py::gen_checkpoint: the reference's PONITA_DIFFUSION
state_dict + hyper_parameters {args Namespace, the reference's AtomicNumberTable instance}); the orientation grid the
reference model was built with travels separately (quirk B2).  The loaded model must reproduce the live reference's
predict_scores on the stored input.  The checkpoint's z_table has 21 states (20 elements + mask): also the
Z != 90 case of the fp16 read-out."""
import numpy as np
import pytest
import torch

from conftest import GOLD, TOL_FP16_MODEL, TOL_FP32, rel_err

pytestmark = pytest.mark.gpu
CKPT = GOLD + "/reference_model.ckpt"


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_reference_pickled_checkpoint_reproduces_reference_outputs(device, gold, precision):
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
    from arreau_b200.tools.atomic_number_table import AtomicNumberTable
    io = gold("reference_model_io.npz")
    m = PONITA_DIFFUSION.load_from_checkpoint(CKPT, ori_grid=io["ori_grid"], strict=True, precision=precision).to(device)
    assert isinstance(m.hparams.z_table, AtomicNumberTable) and list(m.hparams.z_table.zs) == io["zs"].tolist()
    assert m.z_table_zs.tolist() == io["zs"].tolist() and m.diffusion_loss.T == 100
    Zc = len(io["zs"])
    N = io["frac"].shape[0]
    dev = device
    t = torch.full((N,), int(io["timestep"]), device=dev)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(io["types"], device=dev), Zc)
    score, logits, len0 = m.diffusion_loss.predict_scores(
        torch.as_tensor(io["frac"], device=dev), onehot, t, torch.as_tensor(io["num_atoms"], device=dev),
        torch.as_tensor(io["lengths"], device=dev), torch.as_tensor(io["angles"], device=dev), m, None, m.t_emb)
    tol = TOL_FP32 if precision == "fp32" else TOL_FP16_MODEL
    assert rel_err(score.cpu().numpy(), io["score"]) < tol
    assert rel_err(logits.cpu().numpy(), io["logits"]) < tol
    assert rel_err(len0.cpu().numpy(), io["len0"]) < tol
    if precision == "fp16":      # Z + 6 != 96: the per-layer read-out, not the pooled kernels
        assert m.diffusion_loss._engine.pool is None


def test_checkpoint_without_grid_warns_and_strict_rejects_a_misfit(device, gold, tmp_path):
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION, load_checkpoint_file
    with pytest.warns(UserWarning, match="orientation grid"):
        PONITA_DIFFUSION.load_from_checkpoint(CKPT)
    ck = load_checkpoint_file(CKPT)
    del ck["state_dict"]["model.x_embedder.weight"]
    ck["state_dict"]["model.basis_fn.3.bias"] = ck["state_dict"]["model.basis_fn.3.bias"][:-1]
    bad = str(tmp_path / "bad.ckpt")
    torch.save(ck, bad)
    io = gold("reference_model_io.npz")
    with pytest.raises(KeyError, match="x_embedder"):
        PONITA_DIFFUSION.load_from_checkpoint(bad, ori_grid=io["ori_grid"])
    with pytest.warns(UserWarning, match="random initialisation"):
        PONITA_DIFFUSION.load_from_checkpoint(bad, ori_grid=io["ori_grid"], strict=False)


def test_sampling_engine_sees_weights_changed_by_training(device, weights_npz):
    """sample -> train -> sample in one process must not sample from the pre-training weights (the engine holds
    a packed copy)."""
    import argparse
    from arreau_b200.diffusion.lattice_helpers import lattice_from_params
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
    from arreau_b200.synthetic import make_training_batch
    from arreau_b200.tools.atomic_number_table import AtomicNumberTable
    args = argparse.Namespace(dataset="synthetic", lr=1e-2, weight_decay=0.0, epochs=4, warmup=0, layer_scale=1e-6,
                              train_augm=False, hidden_dim=128, layers=5, radius=5.0, num_ori=16, basis_dim=256, degree=3,
                              widening_factor=4, multiple_readouts=True, num_timesteps=1000, max_neighbors=8)
    m = PONITA_DIFFUSION(args, AtomicNumberTable(list(range(1, 90)) + [2001]), ori_grid=weights_npz["ori_grid"])
    m.model.load_state_dict({k: torch.as_tensor(weights_npz[k]) for k in weights_npz.files if k not in ("ori_grid", "fourier_w")})
    m = m.to(device)
    opt = m.configure_optimizers(device)
    cr = make_training_batch(12, seed=5)
    N = cr.total_atoms
    dev = device
    state = (torch.as_tensor(cr.frac, device=dev), torch.nn.functional.one_hot(torch.as_tensor(cr.types, device=dev), 90),
             torch.full((N,), 400, device=dev), torch.as_tensor(cr.num_atoms, device=dev),
             torch.as_tensor(cr.lengths, device=dev), torch.as_tensor(cr.angles, device=dev))
    before = [x.clone() for x in m.diffusion_loss.predict_scores(*state, m, None, m.t_emb)]
    batch = argparse.Namespace(X0=state[0], A0=torch.as_tensor(cr.types, device=dev),
                               L0=lattice_from_params(state[4], state[5]).reshape(-1, 3), num_atoms=state[3])
    for _ in range(3):
        loss = m.training_step(batch)
        loss.backward()
        opt.step()
    after = m.diffusion_loss.predict_scores(*state, m, None, m.t_emb)
    assert float((after[1] - before[1]).abs().max()) > 1e-4          # the new weights are in use
    # and equal to a freshly packed model with the trained parameters
    m.model._packed = None
    m.diffusion_loss._engine = None
    fresh = m.diffusion_loss.predict_scores(*state, m, None, m.t_emb)
    for a, b in zip(after, fresh):
        assert torch.equal(a, b)
