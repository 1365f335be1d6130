import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arreau_b200.diffusion.lattice_dataset import CrystalDataset, save_dataset_npz, batches
from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
from arreau_b200.synthetic import make_training_batch
from arreau_b200.train import default_args
n_cryst, batch = int(sys.argv[1]), int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
dev = torch.device("cuda")
cr = make_training_batch(n_cryst, seed=5)
off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
zs = [cr.types[off[i]:off[i + 1]] % 89 + 1 for i in range(n_cryst)]
frac = [cr.frac[off[i]:off[i + 1]] for i in range(n_cryst)]
lat = np.stack([np.diag(cr.lengths[i]) for i in range(n_cryst)])
ds = CrystalDataset([save_dataset_npz(os.path.join(tempfile.mkdtemp(), "s"), zs, lat, frac)])
torch.manual_seed(0)
model = PONITA_DIFFUSION(default_args(lr=3e-4, epochs=1, warmup=0, batch_size=batch), ds.z_table)
model.to(dev)
model.diffusion_loss.backward_precision = prec
opt = model.configure_optimizers(dev)
for i, b in enumerate(batches(ds, batch, shuffle=True, seed=0, device=dev, rank=0, world=1)):
    loss = model.training_step(b)
    te = model.diffusion_loss.train_engine_for(model.model, model.t_emb, b.num_atoms, dev)
    e = te.eng
    torch.cuda.synchronize()
    print("batch", i, "G", e.G, "N", e.N, "N_cap", e.N_cap, "E", e.num_edges(), "edge_capacity", e.edge_capacity, "loss", float(loss),
          "max atoms", int(b.num_atoms.max()))
    for l in range(2):
        print("  layer", l, "h", float(e.h_debug[l].std()), "x1", float(e.x1_debug[l].std()), "x2", float(e.x2_debug[l].std()),
              "kern", float(e.kernels[l][: e.num_edges()].float().std()), "h nan", bool(torch.isnan(e.h_debug[l]).any()))
    if i >= 1:
        break
