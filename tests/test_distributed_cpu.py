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
