"""Training throughput through the PUBLIC API (`train.fit`: dataset reader -> shuffled batches -> collate -> H2D ->
DiffusionLoss.__call__ -> backward -> fused Adam) on a synthetic Alexandria-shaped dataset written to the .npz twin of
the reference's HDF5 format: python scratch/fit_throughput.py [crystals] [batch] [epochs]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arreau_b200.diffusion.lattice_dataset import CrystalDataset, save_dataset_npz  # noqa: E402
from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION  # noqa: E402
from arreau_b200.synthetic import make_training_batch  # noqa: E402
from arreau_b200.train import default_args, fit  # noqa: E402

n_cryst = int(sys.argv[1]) if len(sys.argv) > 1 else 8100
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 270
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
cr = make_training_batch(n_cryst, seed=5)
off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
zs = [cr.types[off[i]:off[i + 1]] % 89 + 1 for i in range(n_cryst)]
frac = [cr.frac[off[i]:off[i + 1]] for i in range(n_cryst)]
lat = np.stack([np.diag(cr.lengths[i]) for i in range(n_cryst)])
path = save_dataset_npz(os.path.join(tempfile.mkdtemp(), "synthetic"), zs, lat, frac)
t0 = time.time()
ds = CrystalDataset([path])
print(f"dataset: {len(ds)} crystals, {int(cr.num_atoms.sum())} atoms, {len(ds.z_table)} states, loaded in {time.time() - t0:.1f} s", flush=True)
torch.manual_seed(0)
model = PONITA_DIFFUSION(default_args(lr=3e-4, epochs=epochs, warmup=0, batch_size=batch), ds.z_table)
marks = []


def log(msg):
    torch.cuda.synchronize()
    marks.append(time.time())
    print(msg, flush=True)


torch.cuda.synchronize()
marks.append(time.time())
fit(model, ds, epochs, batch, dev, backward_precision="tf32", log=log)
steps = (n_cryst + batch - 1) // batch
for e in range(epochs):
    dt = marks[e + 1] - marks[e]
    print(f"epoch {e}: {dt:.3f} s = {dt / steps * 1e3:.2f} ms per step, {n_cryst / dt:.0f} crystals/s")
dl = model.diffusion_loss
print("train engines built:", dl.train_engine_builds)
