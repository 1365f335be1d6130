"""cProfile of one steady-state epoch of train.fit (host side of the public training API)."""
import cProfile, os, pstats, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arreau_b200.diffusion.lattice_dataset import CrystalDataset, save_dataset_npz
from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
from arreau_b200.synthetic import make_training_batch
from arreau_b200.train import default_args, fit
n_cryst, batch = 8100, 270
dev = torch.device("cuda")
cr = make_training_batch(n_cryst, seed=5)
off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
zs = [cr.types[off[i]:off[i + 1]] % 89 + 1 for i in range(n_cryst)]
frac = [cr.frac[off[i]:off[i + 1]] for i in range(n_cryst)]
lat = np.stack([np.diag(cr.lengths[i]) for i in range(n_cryst)])
ds = CrystalDataset([save_dataset_npz(os.path.join(tempfile.mkdtemp(), "s"), zs, lat, frac)])
torch.manual_seed(0)
model = PONITA_DIFFUSION(default_args(lr=3e-4, epochs=3, warmup=0, batch_size=batch), ds.z_table)
fit(model, ds, 1, batch, dev, backward_precision="tf32", log=lambda *_: None)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
fit(model, ds, 1, batch, dev, backward_precision="tf32", calibrate=False, log=lambda *_: None)
pr.disable()
t_host = time.time() - t0
torch.cuda.synchronize()
print(f"host time of the epoch (30 steps) {t_host * 1e3:.1f} ms, with the GPU drained {(time.time() - t0) * 1e3:.1f} ms")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
