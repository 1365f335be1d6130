"""Host-side logic and the C-ABI boundary, without a GPU: the library loads and exports every symbol the
public header declares, argument validation returns the documented error codes, and the host tables / weight
packing agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, f64_default

HEADER = os.path.join(ROOT, "include", "arreau_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(arreau_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from arreau_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names            # and nothing undeclared is bound
    assert lib.arreau_abi_version() == 1
    dims = [C.c_int() for _ in range(5)]
    assert lib.arreau_model_dims(*[C.byref(d) for d in dims]) == 0
    assert [d.value for d in dims] == [16, 128, 256, 4, 5]


def test_argument_validation_error_codes():
    """Bad arguments are rejected before anything is launched (no GPU needed): negative codes, no exceptions."""
    from arreau_b200 import _lib
    lib = _lib.load()
    assert lib.arreau_graph_scan(None, None, 4, None) == -4                       # ARREAU_ERR_NULL
    assert lib.arreau_lattice_from_params(None, None, 3, None, None) == -4
    assert lib.arreau_lattice_from_params(None, None, 0, None, None) == 0          # empty input is a no-op
    assert lib.arreau_frac_to_cart(None, None, None, 0, None, None) == 0
    buf = (C.c_double * 16)()
    p = C.cast(buf, C.c_void_p)
    assert lib.arreau_lattice_from_params(p, p, -1, p, None) == -1                # ARREAU_ERR_BAD_SHAPE
    assert lib.arreau_graph_count(p, p, p, p, 4, 1, 25.0, 100000, 1, p, p, p, None) == -2   # cap unsupported
    assert lib.arreau_d3pm_reverse(p, p, p, None, 5, p, p, 0.98, 0.02, 1000, 4, 500, p, None) == -1   # Z > 128
    with pytest.raises(RuntimeError, match="ARREAU_ERR_NULL"):
        _lib.call("arreau_graph_scan", None, None, 4, None)
    assert lib.arreau_ponita_forward(None, None, 0, *([None] * 9), 4, 1, 5.0, None, None, None, None) == -4


def test_struct_layouts_match_header():
    """ctypes mirrors of the three argument structs have the field order of the header."""
    from arreau_b200 import _lib
    text = open(HEADER).read()
    for cname, cls in (("arreau_weights", _lib.Weights), ("arreau_workspace", _lib.Workspace), ("arreau_step_args", _lib.StepArgs),
                       ("arreau_step_replay", _lib.StepReplay), ("arreau_workspace_sizes", _lib.WorkspaceSizes)):
        body = text[text.index("typedef struct " + cname):]
        body = body[:body.index("} " + cname)]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split("{", 1)[1].split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = decl.split(",")
            fields.append(re.findall(r"(\w+)\s*$", parts[0])[0])
            fields += [re.findall(r"(\w+)\s*$", p)[0] for p in parts[1:]]
        assert fields == [f[0] for f in cls._fields_], cname


def test_tables_match_golden(gold):
    from arreau_b200.tables import build_tables
    k = gold("kat.npz")
    t = build_tables(1000, 90)
    assert np.array_equal(t.ve_sigmas.numpy(), k["ve_sigmas"]) and np.array_equal(t.vp_betas.numpy(), k["vp_betas"])
    assert np.array_equal(t.vp_alpha_bars.numpy(), k["vp_alpha_bars"]) and np.array_equal(t.vp_sigmas.numpy(), k["vp_sigmas"])
    assert np.array_equal(t.q_keep.numpy(), k["d3pm_keep"]) and np.array_equal(t.q_to_mask.numpy(), k["d3pm_to_mask"])
    assert (t.onestep_keep, t.onestep_to_mask) == (0.98, 0.02)
    assert torch.get_default_dtype() == torch.float32          # the builder restores the caller's default dtype
    # reverse_given_x0 coefficients against the oracle's formula at a few timesteps
    from oracle import restatement as R
    with f64_default():
        tabs = R.DiffusionTables.build(1000, 90)
        xt, x0, z = torch.rand(4, 3) + 5, torch.rand(4, 3) + 5, torch.randn(4, 3)
        for ts in (999, 500, 2, 1):
            ref = R.vp_lattice_reverse_given_x0(tabs, xt, x0, torch.tensor([ts]), z)
            mine = (t.vp_cx0[ts] * x0 + t.vp_cxt[ts] * xt) / t.vp_denom[ts] + t.vp_var[ts] * (z if ts > 1 else 0 * z)
            assert torch.equal(ref, mine), ts


def test_monomial_fold_is_exact():
    """W1 . PolynomialFeatures(3)(v) == W1m . monomials83(v): the fold only merges equal monomials."""
    from arreau_b200.weights import monomial_fold_table
    from oracle import restatement as R
    fold = monomial_fold_table()
    assert fold.shape == (258,) and fold.max() == 82 and len(set(fold.tolist())) == 83
    g = torch.Generator().manual_seed(0)
    with f64_default():
        v = torch.randn(7, 6, generator=g, dtype=torch.float64)
        W = torch.randn(11, 258, generator=g, dtype=torch.float64)
        poly = R.polynomial_features(v, 3)
        mono = torch.zeros(7, 83, dtype=torch.float64)
        mono[:, torch.as_tensor(fold)] = poly                  # equal monomials carry equal values
        Wm = torch.zeros(11, 83, dtype=torch.float64).index_add_(1, torch.as_tensor(fold), W)
        assert torch.allclose(poly @ W.T, mono @ Wm.T, rtol=1e-12, atol=1e-12)
    # order used by the kernels: singles, pairs i<=j, triples i<=j<=k
    assert fold[:6].tolist() == list(range(6)) and fold[6] == 6 and fold[6 + 1] == 7 and fold[6 + 6] == 7


def test_umma_tile_image_layout():
    """16-byte chunk c of row r lands at chunk c ^ (r & 7) of the row's 128 bytes, K slabs back to back."""
    from arreau_b200.weights import umma_tile_image
    rows, K = 16, 128
    w = np.arange(rows * K, dtype=np.float32).reshape(rows, K) % 251        # fp16-exact small integers
    img = umma_tile_image(w).view(torch.float16).float().numpy().reshape(K // 64, rows, 8, 8)
    for r in (0, 1, 7, 9, 15):
        for k in (0, 8, 63, 64, 127):
            slab, c, e = k // 64, (k % 64) // 8, k % 8
            assert img[slab, r, c ^ (r & 7), e] == w[r, k]


def test_angle_factors_reproduce_reference_lattice(gold):
    from arreau_b200.engine import angle_factors
    k = gold("kat.npz")
    f = angle_factors(torch.as_tensor(k["lat_angles"]))
    a, b, c = torch.as_tensor(k["lat_lengths"]).unbind(-1)
    lat = torch.zeros(3, 3, 3, dtype=torch.float64)
    lat[:, 0, 0], lat[:, 0, 2] = a * f[:, 0], a * f[:, 1]
    lat[:, 1, 0], lat[:, 1, 1], lat[:, 1, 2] = -b * f[:, 2] * f[:, 3], b * f[:, 2] * f[:, 4], b * f[:, 5]
    lat[:, 2, 2] = c
    assert np.array_equal(lat.numpy(), k["lat_matrix"])       # bit identical: same torch ops, same order


def test_synthetic_crystals_are_deterministic():
    from arreau_b200.synthetic import CONFIGS, make_crystals
    a, b = make_crystals(16, 2, 20, seed=0), make_crystals(16, 2, 20, seed=0)
    assert np.array_equal(a.frac, b.frac) and np.array_equal(a.num_atoms, b.num_atoms)
    assert 2 <= a.num_atoms.min() and a.num_atoms.max() <= 20 and a.total_atoms == a.frac.shape[0]
    vol = np.abs(np.linalg.det(np.stack([np.diag(l) for l in a.lengths])))
    assert 10 < (vol / a.num_atoms).mean() < 30               # ~18 A^3 per atom (Alexandria density)
    assert set(CONFIGS) >= {"C1", "C2", "C3"}


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under arreau_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "arreau_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_flat_param_layout_matches_reference_state_dict(weights_npz):
    """arreau_train_layout (host-only entry point): every trainable tensor of the reference's state_dict has a slice
    of the flat buffer with its own shape; 1 170 646 parameters in total."""
    import torch
    from arreau_b200.training import FlatParams
    p = FlatParams(164, 4, 90, "cpu")
    names = [k for k in weights_npz.files if k not in ("ori_grid", "fourier_w")]
    assert p.total == 1170646 == sum(weights_npz[k].size for k in names)
    assert set(p.specs) == set(names)
    for k in names:
        assert tuple(p.specs[k][1]) == tuple(weights_npz[k].shape), k
    spans = sorted((off, off + int(np.prod(shape))) for off, shape in p.specs.values())
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == p.total
    p.load_state_dict({k: weights_npz[k] for k in names})
    sd = p.state_dict()
    assert all(np.array_equal(sd[k].numpy(), weights_npz[k]) for k in names)
    # weight-decay groups of lightning_wrappers/diffusion.py:161-186
    mask = p.decay_mask()
    v = p._views(mask)
    assert all(bool(v[k].all()) == (k.endswith(".weight") and ".norm." not in k) for k in names)
    assert all(bool(v[k].any()) == bool(v[k].all()) for k in names)


def test_cosine_warmup_factor():
    from arreau_b200.training import cosine_warmup_factor
    assert abs(cosine_warmup_factor(0, 10, 100)) < 1e-6
    assert abs(cosine_warmup_factor(10, 10, 100) - 0.5 * (1 + np.cos(np.pi * 0.1))) < 1e-12
    assert abs(cosine_warmup_factor(50, 10, 100) - 0.5) < 1e-12


def _ref_args(T=1000):
    import argparse
    return argparse.Namespace(dataset="synthetic", lr=3e-4, weight_decay=0.0, epochs=1, warmup=0, layer_scale=1e-6,
                              train_augm=False, hidden_dim=128, layers=5, radius=5.0, num_ori=16, basis_dim=256, degree=3,
                              widening_factor=4, multiple_readouts=True, num_timesteps=T, max_neighbors=8)


def test_sample_result_writer_round_trip(tmp_path):
    """Wire format of generated crystals (inference/process_generated_crystals.py:8-31): five arrays under crystals/."""
    from arreau_b200.diffusion.diffusion_loss import SampleResult
    from arreau_b200.inference.process_generated_crystals import (KEYS, get_crystal_indexes, load_sample_results_from_hdf5,
                                                                  save_sample_results_to_hdf5)
    rng = np.random.default_rng(0)
    res = SampleResult(frac_x=rng.random((12, 3)), atomic_numbers=rng.integers(1, 90, 12), lattice=rng.random((3, 3, 3)),
                       idx_start=np.array([0, 4, 8]), num_atoms=np.array([4, 4, 4]))
    path = save_sample_results_to_hdf5(res, str(tmp_path / "out" / "crystals.h5"))
    back = load_sample_results_from_hdf5(path)
    for k in KEYS:
        assert np.array_equal(getattr(back, k), getattr(res, k)), k
    assert get_crystal_indexes(back, 1) == (4, 8)


def test_checkpoint_ingestion_remaps_reference_classes(tmp_path, weights_npz):
    """A Lightning-shaped .ckpt whose pickled z_table lives at the REFERENCE's module path
    (diffusion.tools.atomic_number_table) loads into the mirror by parameter name."""
    import sys
    import types
    import torch
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION
    mod_names = ["diffusion", "diffusion.tools", "diffusion.tools.atomic_number_table"]
    saved = {n: sys.modules.get(n) for n in mod_names}
    fake = types.ModuleType("diffusion.tools.atomic_number_table")

    class AtomicNumberTable:   # stands in for the reference's class at pickling time
        def __init__(self, zs):
            self.zs = zs
    AtomicNumberTable.__module__ = "diffusion.tools.atomic_number_table"
    AtomicNumberTable.__qualname__ = "AtomicNumberTable"
    fake.AtomicNumberTable = AtomicNumberTable
    for n in mod_names[:2]:
        sys.modules[n] = types.ModuleType(n)
    sys.modules[mod_names[2]] = fake
    try:
        sd = {"model." + k: torch.as_tensor(weights_npz[k]) for k in weights_npz.files if k not in ("ori_grid", "fourier_w")}
        sd["t_emb.gaussian_fourier_proj_w"] = torch.as_tensor(weights_npz["fourier_w"])
        sd["z_table_zs"] = torch.tensor(list(range(1, 90)) + [2001])
        sd["diffusion_loss.d3pm.q_mats"] = torch.zeros(2, 2)          # ignored: tables are rebuilt
        ckpt = {"state_dict": sd, "hyper_parameters": {"args": _ref_args(), "z_table": AtomicNumberTable(list(range(1, 90)) + [2001])}}
        path = str(tmp_path / "model.ckpt")
        torch.save(ckpt, path)
    finally:
        for n in mod_names:
            if saved[n] is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = saved[n]
    m = PONITA_DIFFUSION.load_from_checkpoint(path, ori_grid=weights_npz["ori_grid"], strict=True)
    own = m.state_dict()
    for k in weights_npz.files:
        if k not in ("ori_grid", "fourier_w"):
            assert np.array_equal(own["model." + k].numpy().astype(np.float32), weights_npz[k]), k
    assert np.array_equal(own["t_emb.gaussian_fourier_proj_w"].numpy().astype(np.float32), weights_npz["fourier_w"])
    assert m.z_table_zs.tolist()[-1] == 2001 and len(m.hparams.z_table) == 90
    # round trip through the mirror's own checkpoint writer (carries the orientation grid, quirk B2)
    p2 = str(tmp_path / "m2.ckpt")
    m.save_checkpoint(p2)
    m2 = PONITA_DIFFUSION.load_from_checkpoint(p2, strict=True)
    assert np.allclose(m2.model.ori_grid.numpy(), weights_npz["ori_grid"])


def test_atomic_symbols_to_indices():
    from arreau_b200.tools.atomic_number_table import AtomicNumberTable, atomic_symbols_to_indices
    zt = AtomicNumberTable(list(range(1, 90)) + [2001])
    assert atomic_symbols_to_indices(zt, ["H", "C", "Ac"]).tolist() == [0, 5, 88]


def test_crystal_dataset_and_collate(tmp_path):
    """Input pipeline (lattice_dataset.py): dataset file -> items -> flat batch, sharded over 2 ranks."""
    from arreau_b200.diffusion.lattice_dataset import CrystalDataset, batches, collate_crystals, save_dataset_npz
    from arreau_b200.synthetic import make_crystals
    cr = make_crystals(7, 1, 9, seed=2)
    off = np.concatenate([[0], np.cumsum(cr.num_atoms)])
    zs = [cr.types[off[i]:off[i + 1]] + 1 for i in range(7)]                    # atomic numbers 1..89
    frac = [cr.frac[off[i]:off[i + 1]] for i in range(7)]
    lat = np.stack([np.diag(cr.lengths[i]) for i in range(7)])
    path = save_dataset_npz(str(tmp_path / "alexandria_tiny"), zs, lat, frac)
    ds = CrystalDataset([path])
    assert len(ds) == 7 and ds.z_table.zs[-1] == 2001 and len(ds.z_table) == len(set(np.concatenate(zs))) + 1
    b = collate_crystals([ds[i] for i in range(7)])
    assert b.X0.shape == (cr.total_atoms, 3) and b.L0.shape == (21, 3) and b.num_atoms.tolist() == cr.num_atoms.tolist()
    assert np.array_equal(b.batch.numpy(), np.repeat(np.arange(7), cr.num_atoms))
    assert np.array_equal(np.asarray(ds.z_table.zs)[b.A0.numpy()], np.concatenate(zs))
    assert np.allclose(b.pos.numpy(), cr.frac * cr.lengths[b.batch.numpy()])
    seen = []
    for rank in range(2):
        for bb in batches(ds, 2, shuffle=True, seed=3, rank=rank, world=2):
            seen.append(int(bb.num_atoms.shape[0]))
    assert sum(seen) == 7


def test_every_rank_gets_the_same_number_of_batches():
    """each training step carries a gradient all-reduce, so the ranks must run the same number of
    steps for ANY dataset size: the batch list is padded to a multiple of the world size by wrapping around (torch's
    DistributedSampler, drop_last=False).  Before the fix 7 crystals / batch 3 / world 2 gave 2 and 1 batches."""
    from arreau_b200.diffusion.lattice_dataset import batch_index_lists
    for n, bs, world in [(7, 3, 2), (7, 2, 2), (10, 3, 4), (1, 4, 8), (270 * 5 + 1, 270, 8), (9, 3, 3), (0, 4, 2)]:
        per_rank = [batch_index_lists(n, bs, True, 5, r, world) for r in range(world)]
        counts = [len(p) for p in per_rank]
        assert len(set(counts)) == 1, (n, bs, world, counts)
        covered = set(int(i) for p in per_rank for b in p for i in b)
        assert covered == set(range(n)), (n, bs, world)                      # nobody's crystals are dropped
        assert sum(len(b) for p in per_rank for b in p) >= n
        # the padding re-uses whole leading batches, at most world - 1 of them
        assert sum(counts) - (n + bs - 1) // bs < world if n else counts == [0] * world


def test_reference_pickled_checkpoint_loads_on_the_host(gold):
    """tests/golden/reference_model.ckpt was pickled by the LIVE reference classes (oracle/gen_golden.py::gen_checkpoint):
    hyper_parameters.z_table is a `diffusion.tools.atomic_number_table.AtomicNumberTable` in the pickle and resolves to
    this package's class; every model / time-embedding tensor lands under the same name (no compute: CPU only)."""
    import zipfile
    from conftest import GOLD
    from arreau_b200.lightning_wrappers.diffusion import PONITA_DIFFUSION, load_checkpoint_file
    from arreau_b200.tools.atomic_number_table import AtomicNumberTable
    path = GOLD + "/reference_model.ckpt"
    with zipfile.ZipFile(path) as z:
        pkl = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    assert b"diffusion.tools.atomic_number_table" in pkl                     # the reference's own class path
    ck = load_checkpoint_file(path)
    assert isinstance(ck["hyper_parameters"]["z_table"], AtomicNumberTable)
    io = gold("reference_model_io.npz")
    m = PONITA_DIFFUSION.load_from_checkpoint(path, ori_grid=io["ori_grid"], strict=True)
    own = m.state_dict()
    n_model = 0
    for k, v in ck["state_dict"].items():
        if (k.startswith("model.") or k.startswith("t_emb.")) and k in own and v.numel():
            assert torch.equal(own[k].float().cpu(), v.float()), k
            n_model += v.numel()
    assert n_model > 1_100_000 and m.diffusion_loss.num_atomic_states == 21 and m.diffusion_loss.T == 100


def test_workspace_bytes_query():
    """arreau_workspace_bytes (the caller sizes and owns every buffer) against the sizes the engine allocates,
    for both precision paths (host-only: no kernel is launched)."""
    from arreau_b200 import _lib
    lib = _lib.load()
    N, G, cap, F, Z = 1000, 37, 8, 164, 90
    for prec, kbytes in ((_lib.PRECISION_FP32, 4), (_lib.PRECISION_FP16, 2)):
        o = _lib.WorkspaceSizes()
        assert lib.arreau_workspace_bytes(N, G, N * cap, prec, F, Z, C.byref(o)) == 0
        node = N * 16 * 128
        assert o.h == node * 4 and o.x1 == node * 4 and o.debug_per_layer == node * 4
        assert o.kernels == 5 * N * cap * 16 * 128 * kbytes
        assert o.y == (node * 4 if kbytes == 4 else ((N * 16 + 127) // 128) * 128 * 128 * 2)
        assert o.acc == N * 96 * 4 and o.pool == (0 if kbytes == 4 else 6 * ((N + 15) // 16) * 4 * 128 * 16 * 4)
        assert (o.x, o.vec, o.logits, o.score, o.len0) == (N * F * 4, N * 48, N * Z * 4, N * 12, G * 12)
        assert (o.pos, o.raw_count, o.deg, o.row_ptr, o.num_neighbors_image) == (N * 24, N * 4, N * 4, (N + 1) * 4, G * 8)
        assert (o.src, o.dst, o.cell, o.dist, o.dir) == (N * cap * 4, N * cap * 4, N * cap, N * cap * 8, N * cap * 24)
        assert (o.z_len, o.z_frac, o.u_type) == (G * 24, N * 24, N * Z * 8)
    o = _lib.WorkspaceSizes()
    assert lib.arreau_workspace_bytes(10, 2, 80, _lib.PRECISION_FP16, 95, 21, C.byref(o)) == 0 and o.pool == 0   # Z != 90
    assert lib.arreau_workspace_bytes(-1, 2, 80, 0, 164, 90, C.byref(o)) == -1
    assert lib.arreau_workspace_bytes(10, 2, 80, 2, 164, 90, C.byref(o)) == -2


def test_device_names_compare_like_torch_places_tensors():
    """DiffusionLoss caches engines per parameter buffer and device: "cuda" (the current device) and "cuda:N" must not be
    taken for different places (it would re-flatten the parameters under the optimizer)."""
    from arreau_b200.diffusion.diffusion_loss import DiffusionLoss
    same = DiffusionLoss._same_device
    assert same("cpu", torch.device("cpu")) and not same("cpu", "cuda") and not same("cuda:0", "cpu")
    assert same("cuda:1", torch.device("cuda", 1)) and not same("cuda:0", "cuda:1")


def test_vectorised_collate_matches_the_per_item_collate(tmp_path):
    """CrystalDataset.collate_indices (what the epoch iterator uses: gathers over a flat copy of the set) returns the
    same batch as collate_crystals over the items, repeated and out-of-order indices included, plus the host copy of the
    atoms-per-crystal vector the engines bind the topology from."""
    from arreau_b200.diffusion.lattice_dataset import CrystalDataset, batches, collate_crystals, save_dataset_npz
    rng = np.random.default_rng(3)
    n = 23
    na = rng.integers(1, 9, size=n)
    zs = [rng.integers(1, 20, size=k) for k in na]
    frac = [rng.random((k, 3)) for k in na]
    lat = rng.random((n, 3, 3)) + np.eye(3) * 3
    ds = CrystalDataset([save_dataset_npz(str(tmp_path / "d"), zs, lat, frac)])
    idx = [5, 0, 22, 7, 7, 3]
    a, b = collate_crystals([ds[i] for i in idx]), ds.collate_indices(idx)
    for k in ("X0", "A0", "L0", "num_atoms", "batch", "pos"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(b.num_atoms_cpu, na[idx]) and np.array_equal(a.num_atoms_cpu, na[idx])
    seen = np.concatenate([x.num_atoms_cpu for x in batches(ds, 5, shuffle=True, seed=2)])
    assert sorted(seen.tolist()) == sorted(na.tolist())
