"""Multi-rank host logic on CPU (gloo, world_size 2): crystal sharding and the final gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from arreau_b200.diffusion.diffusion_loss import SampleResult
    from arreau_b200.distributed import gather_sample_results, shard_range
    G = 7
    sizes = np.array([3, 5, 2, 4, 6, 1, 3])
    lo, hi = shard_range(G, rank, world)
    na = sizes[lo:hi]
    start = int(sizes[:lo].sum())
    n = int(na.sum())
    local = SampleResult(frac_x=(np.arange(start, start + n)[:, None] + np.array([0.1, 0.2, 0.3])[None]),
                         atomic_numbers=np.arange(start, start + n) % 89 + 1,
                         lattice=np.arange(lo, hi)[:, None, None] * np.ones((1, 3, 3)), num_atoms=na)
    full = gather_sample_results(local)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), frac=full.frac_x, z=full.atomic_numbers, lat=full.lattice,
             na=full.num_atoms, idx=full.idx_start)
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    from arreau_b200.distributed import shard_range
    assert [shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_range(65536, r, 8) for r in range(8)][-1] == (57344, 65536)
    cover = [shard_range(10, r, 4) for r in range(4)]
    assert cover[0][0] == 0 and cover[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sizes = np.array([3, 5, 2, 4, 6, 1, 3])
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["na"], sizes)
        assert np.array_equal(z["idx"], np.concatenate([[0], np.cumsum(sizes)[:-1]]))
        assert np.allclose(z["frac"][:, 0], np.arange(sizes.sum()) + 0.1)
        assert np.array_equal(z["z"], np.arange(sizes.sum()) % 89 + 1)
        assert np.array_equal(z["lat"][:, 0, 0], np.arange(7))


def _ddp_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from arreau_b200.distributed import allreduce_gradients, broadcast_parameters, reduce_loss_metric
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        grad = torch.full((1170646,), float(rank + 1))
        allreduce_gradients(grad)
        params = torch.full((16,), float(rank))
        broadcast_parameters(params, src=0)
        metric = reduce_loss_metric(torch.tensor(3.0 * (rank + 1)), torch.tensor(rank + 1))
        out[rank] = (float(grad[0]), float(grad[-1]), float(params.sum()), float(metric))
    finally:
        dist.destroy_process_group()


def test_ddp_gradient_allreduce_world2_gloo():
    """C5's only collectives: mean of the flat gradient buffer, the initial weight broadcast and the
    2-scalar loss metric, over a 2-rank gloo group on CPU."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    for r in range(2):
        g0, g1, psum, metric = out[r]
        assert g0 == g1 == 1.5 and psum == 0.0 and abs(metric - 3.0) < 1e-12


class _StubModel:
    """Stands in for PONITA_DIFFUSION.sample (lightning_wrappers/diffusion.py:221-228) on CPU: crystal k of the
    whole job is recognisable from the seed the driver passes (seed + first crystal of the batch)."""

    def __init__(self):
        self.calls = []

    def sample(self, num_atoms_per_sample, num_samples_in_batch, use_constant_atomic_symbols=None, device=None,
               device_noise=True, seed=0):
        from arreau_b200.diffusion.diffusion_loss import SampleResult
        self.calls.append((seed, num_samples_in_batch))
        n, g = num_atoms_per_sample, num_samples_in_batch
        ids = seed + np.arange(g)                                   # global crystal ids of this batch
        return SampleResult(frac_x=np.repeat(ids, n)[:, None] / 1000.0 + np.zeros((1, 3)),
                            atomic_numbers=np.repeat(ids % 89 + 1, n), lattice=ids[:, None, None] * np.ones((1, 3, 3)),
                            num_atoms=np.full(g, n), idx_start=np.arange(0, g * n, n))


def _generate_worker(rank, world, port, out_dir, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arreau_b200.generate import generate_n_crystals
        m = _StubModel()
        res = generate_n_crystals(m, total, 3, None, num_crystals_per_batch=4, device="cpu", seed=0,
                                  out_path=os.path.join(out_dir, f"crystals_{total}.npz"))
        np.savez(os.path.join(out_dir, f"gen{total}_r{rank}.npz"), frac=res.frac_x, z=res.atomic_numbers, lat=res.lattice,
                 na=res.num_atoms, idx=res.idx_start, calls=np.array(m.calls, dtype=np.int64).reshape(-1, 2))
    finally:
        dist.destroy_process_group()


def test_generate_driver_shards_batches_and_gathers_world2(tmp_path):
    """main_diffusion_generate.py:52-94 over 2 ranks: contiguous crystal blocks per rank, batches of
    at most num_crystals_per_batch inside a block (ragged last batch), one gather, rank 0 writes the file.  Also the
    degenerate job with fewer crystals than ranks (rank 1 owns nothing and still takes part in the gather)."""
    from arreau_b200.inference.process_generated_crystals import load_sample_results_from_hdf5
    for total in (11, 1):
        mp.spawn(_generate_worker, args=(2, _free_port(), str(tmp_path), total), nprocs=2, join=True)
        for r in range(2):
            z = np.load(tmp_path / f"gen{total}_r{r}.npz")
            assert np.array_equal(z["na"], np.full(total, 3)) and np.array_equal(z["idx"], np.arange(0, 3 * total, 3))
            assert np.array_equal(z["lat"][:, 0, 0], np.arange(total))              # crystal order = global order
            assert np.array_equal(z["z"], np.repeat(np.arange(total) % 89 + 1, 3))
            assert np.allclose(z["frac"][:, 0], np.repeat(np.arange(total), 3) / 1000.0)
        calls0 = np.load(tmp_path / f"gen{total}_r0.npz")["calls"].tolist()
        calls1 = np.load(tmp_path / f"gen{total}_r1.npz")["calls"].tolist()
        if total == 11:      # rank 0 owns [0, 6): batches 4 + 2; rank 1 owns [6, 11): 4 + 1
            assert calls0 == [[0, 4], [4, 2]] and calls1 == [[6, 4], [10, 1]]
        else:
            assert calls0 == [[0, 1]] and calls1 == []
        back = load_sample_results_from_hdf5(str(tmp_path / f"crystals_{total}.npz"))
        assert back.num_atoms.shape[0] == total and np.array_equal(back.lattice[:, 0, 0], np.arange(total))
