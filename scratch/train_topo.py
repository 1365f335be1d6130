"""Cost of building a TrainEngine per batch topology (variable training batches)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from arreau_b200.synthetic import make_training_batch
from arreau_b200.tables import build_tables
from arreau_b200.training import FlatParams, TrainEngine
dev = torch.device('cuda')
w = np.load('tests/golden/weights_seed0.npz')
sd = {k: w[k] for k in w.files if k not in ('ori_grid', 'fourier_w')}
p = FlatParams(164, 4, 90, dev); p.load_state_dict(sd)
tabs = build_tables(1000, 90)
for i in range(6):
    cr = make_training_batch(270, seed=i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    te = TrainEngine(p, tabs, w['fourier_w'], w['ori_grid'], cr.num_atoms, 5.0, 8, device=dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f'batch {i}: N={cr.total_atoms} engine build {1e3*(t1-t0):.1f} ms, mem {torch.cuda.memory_allocated()/1e9:.2f} GB')
    del te
