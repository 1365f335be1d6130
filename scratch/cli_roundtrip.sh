#!/bin/bash
# public CLIs end to end: write a small dataset, train one epoch, sample from the checkpoint
set -e
mkdir -p gpurun_out /tmp/rt
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from arreau_b200.diffusion.lattice_dataset import save_dataset_npz
from arreau_b200.synthetic import make_training_batch
cr = make_training_batch(600, seed=3)
off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
zs = [cr.types[off[i]:off[i + 1]] % 30 + 1 for i in range(600)]
frac = [cr.frac[off[i]:off[i + 1]] for i in range(600)]
lat = np.stack([np.diag(cr.lengths[i]) for i in range(600)])
print(save_dataset_npz("/tmp/rt/data", zs, lat, frac))
PY
python -m arreau_b200.train --data /tmp/rt/data.npz --epochs 2 --batch_size 100 --out /tmp/rt/model.ckpt
python -m arreau_b200.generate --help | head -30
python -m arreau_b200.generate --model_path /tmp/rt/model.ckpt --num_crystals 16 --num_atoms 6 --out /tmp/rt/gen 2>&1 | tail -5 || python -m arreau_b200.generate --model_path /tmp/rt/model.ckpt --num_crystals 16 --num_atoms 6 2>&1 | tail -5
ls -la /tmp/rt
