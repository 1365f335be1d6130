#!/bin/bash
# the training CLI under torchrun on 2 GPUs (DDP: equal steps per rank, one gradient all-reduce per step), then sampling from its checkpoint
set -e
mkdir -p /tmp/rt
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from arreau_b200.diffusion.lattice_dataset import save_dataset_npz
from arreau_b200.synthetic import make_training_batch
cr = make_training_batch(1100, seed=4)
off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
zs = [cr.types[off[i]:off[i + 1]] % 30 + 1 for i in range(1100)]
frac = [cr.frac[off[i]:off[i + 1]] for i in range(1100)]
lat = np.stack([np.diag(cr.lengths[i]) for i in range(1100)])
print(save_dataset_npz("/tmp/rt/data2", zs, lat, frac))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 -m arreau_b200.train --data /tmp/rt/data2.npz --epochs 2 --batch_size 100 --out /tmp/rt/model2.ckpt
python -m arreau_b200.generate --model_path /tmp/rt/model2.ckpt --num_crystals 8 --num_atoms 5 --out /tmp/rt/gen2 2>&1 | tail -2
