"""C4 (box-sharded sampling) through the public driver: torchrun --nproc-per-node N scratch/generate_mg.py [crystals_per_gpu]
Every rank runs whole T=1000 trajectories of its own crystals (batches of 1024 x 40 atoms) through
arreau_b200.generate.generate_n_crystals; the only collective is the final NCCL gather; rank 0 writes crystals.h5 and
prints one JSON line (wall clock around the whole call, synchronised on both sides)."""
import argparse, json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arreau_b200.generate import generate_n_crystals
from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
from arreau_b200.synthetic import calibrate_length_readout
from arreau_b200.tools.atomic_number_table import AtomicNumberTable

Z, n = 90, 40
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
local, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_seed0.npz"))
args = argparse.Namespace(dataset="synthetic", lr=1e-3, weight_decay=0.0, epochs=4, warmup=1, layer_scale=1e-6, train_augm=False,
                          hidden_dim=128, layers=5, radius=5.0, num_ori=16, basis_dim=256, degree=3, widening_factor=4,
                          multiple_readouts=True, num_timesteps=1000, max_neighbors=8)
m = PONITA_DIFFUSION(args, AtomicNumberTable(list(range(1, Z)) + [2001]), ori_grid=w["ori_grid"], precision="fp16")
sd = calibrate_length_readout({k: w[k] for k in w.files if k not in ("ori_grid", "fourier_w")}, n)
m.model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
with torch.no_grad():
    m.t_emb.gaussian_fourier_proj_w.copy_(torch.as_tensor(w["fourier_w"]))
m = m.to(dev)
generate_n_crystals(m, 8 * world, n, None, num_crystals_per_batch=8, device=dev)          # warm-up (engine, NCCL)
total = per_gpu * world
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
tm = {}
res = generate_n_crystals(m, total, n, None, num_crystals_per_batch=1024, device=dev, out_path="gpurun_out/crystals_mg.h5",
                          timings=tm)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if local == 0:
    ok = bool(res.num_atoms.shape[0] == total and np.isfinite(res.frac_x).all() and np.isfinite(res.lattice).all()
              and (res.frac_x >= 0).all() and (res.frac_x < 1).all())
    print(json.dumps({"workload": "C4: generate_n_crystals, T=1000, 40 atoms, cap 8, fp16 tensor path",
                      "n_gpus": world, "crystals": total, "crystals_per_gpu": per_gpu, "seconds": dt,
                      "crystals_per_sec": total / dt, "gathered_ok": ok, "atoms": int(res.frac_x.shape[0]),
                      "rank0_seconds": {"sample_batches": [round(x, 3) for x in tm["sample"]], "gather": round(tm["gather"], 4),
                                        "write_file": round(tm["write"], 4)},
                      "unique_types": int(np.unique(res.atomic_numbers).size)}))
    for f in ("gpurun_out/crystals_mg.h5", "gpurun_out/crystals_mg.npz"):
        if os.path.exists(f):
            os.remove(f)
if world > 1:
    dist.destroy_process_group()
