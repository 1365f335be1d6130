"""compute-sanitizer target: one C1-sized denoise step on both precision paths, a ragged batch, an uncapped batch
(long-row message pass) and one training step (fp32 and TF32 backward).  Small on purpose: the sanitizer runs kernels
10-100x slower.   compute-sanitizer --tool memcheck|racecheck|synccheck python scratch/sanitize_step.py [what ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arreau_b200.engine import DenoiseEngine  # noqa: E402
from arreau_b200.synthetic import make_crystals, make_training_batch  # noqa: E402
from arreau_b200.tables import build_tables  # noqa: E402
from arreau_b200.weights import PonitaWeights  # noqa: E402

what = set(sys.argv[1:]) or {"fp16", "fp32", "ragged", "uncapped", "train"}
dev = torch.device("cuda:0")
w = np.load(os.path.join(ROOT, "tests", "golden", "weights_seed0.npz"))
sd = {k: w[k] for k in w.files if k not in ("ori_grid", "fourier_w")}
packed = PonitaWeights(sd, w["ori_grid"], device=dev)
tabs = build_tables(1000, 90)


def step(cr, precision, cap=8, t=400, radius=5.0):
    eng = DenoiseEngine(packed, tabs, w["fourier_w"], cr.num_atoms, radius, cap, precision=precision, device=dev)
    eng.set_state(cr.frac, cr.types, cr.lengths, cr.angles)
    eng.draw_noise(1, 0)
    eng.step(t)
    eng.draw_noise(1, 1)
    eng.step(t - 1)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(eng.score).all())
    print(f"ok {precision} G={cr.num_crystals} N={cr.total_atoms} cap={cap} E={eng.num_edges()}", flush=True)


if "fp16" in what:
    step(make_crystals(16, 2, 20, seed=0), "fp16")
if "fp32" in what:
    step(make_crystals(16, 2, 20, seed=0), "fp32")
if "ragged" in what:
    cr = make_crystals(5, 1, 17, seed=4)
    step(cr, "fp16")
    step(make_crystals(1, 1, None, seed=5), "fp16")
if "uncapped" in what:
    step(make_crystals(3, 12, 30, seed=6), "fp16", cap=0)
if "train" in what:
    from arreau_b200.diffusion.lattice_helpers import lattice_from_params
    from arreau_b200.training import FlatParams, FusedAdam, TrainEngine
    for prec in ("fp32", "tf32"):
        p = FlatParams(164, 4, 90, dev)
        p.load_state_dict(sd)
        cr = make_training_batch(12, seed=2)
        te = TrainEngine(p, tabs, w["fourier_w"], w["ori_grid"], cr.num_atoms, 5.0, 8, device=dev, backward_precision=prec)
        opt = FusedAdam(p, lr=3e-4)
        G, N = cr.num_crystals, cr.total_atoms
        g = torch.Generator(device=dev).manual_seed(3)
        lat0 = lattice_from_params(torch.as_tensor(cr.lengths).to(dev), torch.as_tensor(cr.angles).to(dev))
        te.loss_and_grads(torch.as_tensor(cr.frac).to(dev), torch.as_tensor(cr.types).to(dev), lat0,
                          torch.randint(1, 1001, (G,), device=dev, generator=g),
                          torch.randn(N, 3, device=dev, dtype=torch.float64, generator=g),
                          torch.rand(N, 90, device=dev, dtype=torch.float64, generator=g),
                          torch.randn(G, 3, device=dev, dtype=torch.float64, generator=g))
        opt.step()
        torch.cuda.synchronize()
        print(f"ok train {prec} loss={te.loss.tolist()}", flush=True)
