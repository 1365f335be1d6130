"""Mirror of the training entry point (main_diffusion.py:284-307) without Lightning:

    python -m arreau_b200.train --data out/alexandria_ps_000.npz --epochs 2 --batch_size 270
    torchrun --nproc-per-node 8 -m arreau_b200.train ...        # data parallel: one all-reduce of the flat gradient per step

Per step: DiffusionLoss.__call__ (noising, predict_scores, loss, hand-written backward -> flat gradient buffer),
gradient all-reduce (mean) across ranks, fused Adam with global-norm clip 0.5 and the cosine-warmup schedule per epoch;
the first batch runs the one-time `callibrate` pass (ponita/nn/conv.py:122-123) like the reference's first train-mode
forward.  The loss metric is reduced like DiffusionLossMetric (sum of losses / number of crystals over all ranks)."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch
import torch.distributed as dist

from .diffusion.lattice_dataset import CrystalDataset, batches
from .distributed import allreduce_gradients, broadcast_parameters, reduce_loss_metric
from .lightning_wrappers.diffusion import PONITA_DIFFUSION


def default_args(**over):
    """argparse defaults of main_diffusion.py:88-120 + Makefile:7."""
    a = argparse.Namespace(dataset="alexandria", lr=3e-4, weight_decay=0.0, epochs=1, warmup=0, layer_scale=1e-6,
                           train_augm=False, hidden_dim=128, layers=5, radius=5.0, num_ori=16, basis_dim=256, degree=3,
                           widening_factor=4, multiple_readouts=True, num_timesteps=1000, max_neighbors=8, batch_size=270)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def fit(model: PONITA_DIFFUSION, dataset: CrystalDataset, epochs: int, batch_size: int, device, seed: int = 0,
        backward_precision: str = "tf32", calibrate: bool = True, log=print):
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    model.to(device)
    model.diffusion_loss.backward_precision = backward_precision
    opt = model.configure_optimizers(device)
    flat = model.model.flat
    broadcast_parameters(flat.data)
    history = []
    for epoch in range(epochs):
        opt.set_epoch(epoch, model.warmup, max(model.epochs, 1))
        total_loss = torch.zeros((), dtype=torch.float64, device=device)
        total_samples = torch.zeros((), dtype=torch.float64, device=device)
        for batch in batches(dataset, batch_size, shuffle=True, seed=seed + epoch, device=device, rank=rank, world=world):
            loss = model.training_step(batch)          # the step's kernels have filled flat.grad
            if calibrate:
                # ponita/nn/conv.py:122-123: the reference's FIRST train-mode forward -- the noised batch of the first
                # training step at its random timesteps -- re-initialises the kernel / fiber_kernel scales from its own
                # activations AFTER using the old weights, and that same forward's loss is then back-propagated: the
                # first gradient belongs to the un-rescaled weights and is applied to the rescaled ones.  Same here:
                # statistics from the step that has just run, rescale, then the optimizer step.  (Under DDP every
                # reference rank rescales from its own shard and the replicas silently diverge; here rank 0's scales
                # are broadcast -- the one deliberate deviation.)
                te = model.diffusion_loss.train_engine_for(model.model, model.t_emb, getattr(batch, "num_atoms_cpu", batch.num_atoms), device)
                te.calibrate_from_last_forward()
                broadcast_parameters(flat.data)
                for layer in model.model.interaction_layers:
                    layer.conv.callibrated.fill_(True)
                calibrate = False
            allreduce_gradients(flat.grad)
            opt.step()
            total_loss += loss.detach().double()
            total_samples += batch.num_atoms.shape[0]
        metric = reduce_loss_metric(total_loss, total_samples)
        history.append(float(metric))
        if rank == 0:
            log(f"epoch {epoch}: train loss {history[-1]:.6f} (lr {opt.lr:.3e})")
    return history


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", nargs="+", required=True, help="dataset files (.h5 as written by prep_datasets.py, or .npz)")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batch_size", type=int, default=270)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--backward_precision", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--out", default="out/model.ckpt")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=device)
    ds = CrystalDataset(args.data)
    torch.manual_seed(0)
    model = PONITA_DIFFUSION(default_args(lr=args.lr, epochs=args.epochs, warmup=args.warmup, batch_size=args.batch_size), ds.z_table)
    fit(model, ds, args.epochs, args.batch_size, device, backward_precision=args.backward_precision)
    if not dist.is_initialized() or dist.get_rank() == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)) or ".", exist_ok=True)
        model.save_checkpoint(args.out)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
