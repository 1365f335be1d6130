"""The oracle (oracle/restatement.py, a CPU restatement of the reference step) against the golden vectors
written from the LIVE reference by oracle/gen_golden.py.  CPU only."""
import numpy as np
import torch

from conftest import f64_default
from oracle import restatement as R

T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731


def _weights(w, radius=5.0, sd=None):
    sd = sd if sd is not None else {k: w[k] for k in w.files if k not in ("ori_grid", "fourier_w")}
    return R.PonitaWeights({k: T64(v) for k, v in sd.items()}, T64(w["ori_grid"]), radius)


def test_schedules_and_kats(gold):
    k = gold("kat.npz")
    with f64_default():
        tabs = R.DiffusionTables.build(1000, 90)
        assert np.array_equal(tabs.ve_sigmas.numpy(), k["ve_sigmas"])
        assert tabs.vp_alpha_bars.dtype == torch.float32 and np.array_equal(tabs.vp_alpha_bars.numpy(), k["vp_alpha_bars"])
        assert np.array_equal(tabs.vp_betas.numpy(), k["vp_betas"]) and np.array_equal(tabs.vp_sigmas.numpy(), k["vp_sigmas"])
        assert np.array_equal(tabs.q_mats[:, 0, 0].numpy(), k["d3pm_keep"])
        assert np.array_equal(tabs.q_mats[:, 0, 89].numpy(), k["d3pm_to_mask"])
        assert np.array_equal(tabs.q_one_step_transposed[0].numpy(), k["d3pm_onestep_T"])
        lat = R.lattice_from_params(T64(k["lat_lengths"]), T64(k["lat_angles"]))
        assert np.array_equal(lat.numpy(), k["lat_matrix"])
        assert np.array_equal(R.frac_to_cart_coords(T64(k["f2c_frac"]), lat, torch.tensor([1, 1, 1])).numpy(), k["f2c_cart"])
        assert np.array_equal(R.polynomial_features(T64(k["poly_in"]), 3).numpy(), k["poly_out"])
        assert np.array_equal(R.polynomial_cutoff(T64(k["cut_in"]), 5.0).numpy(), k["cut_out"])
    # SURVEY Appendix C spot values
    assert abs(k["ve_sigmas"][500] - 0.0316227766016838) < 1e-15 and k["ve_sigmas"][1000] == 1.0
    assert abs(k["d3pm_keep"][0] - 0.98) < 1e-15 and abs(k["d3pm_to_mask"][0] - 0.02) < 1e-15


def test_graph_cases(graph_cases):
    with f64_default():
        for c in graph_cases:
            got = R.radius_graph_pbc(T64(c["cart"]), T64(c["lattice"]), torch.as_tensor(c["num_atoms"]), float(c["radius"]), int(c["cap"]))
            name = str(c["name"])
            assert np.array_equal(got[0].numpy(), c["edge_index"]), name
            assert np.array_equal(got[1].numpy(), c["cell_offsets"]), name
            assert np.array_equal(got[2].numpy(), c["num_neighbors_image"]), name
            assert np.array_equal(got[3].numpy(), c["dist"]) and np.array_equal(got[4].numpy(), c["direction"]), name
    names = [str(c["name"]) for c in graph_cases]
    tie = graph_cases[names.index("tie_1atom_cap8")]
    # SURVEY Appendix C, canonical (stable) tie rule: cells [1,3,4,10,12,14,16,22]
    cells = (-tie["cell_offsets"] + 1) @ np.array([9, 3, 1])
    assert cells.astype(int).tolist() == [1, 3, 4, 10, 12, 14, 16, 22]
    assert graph_cases[names.index("close_pair")]["edge_index"].shape[1] == 0


def test_forward_matches_live_reference(gold):
    f, w = gold("forward_c1_t500.npz"), gold("weights_seed0.npz")
    with f64_default():
        logits, vec, len0 = R.ponita_forward(_weights(w), T64(f["x"]), T64(f["vec"]), torch.as_tensor(f["edge_index"]),
                                             T64(f["dist"]), T64(f["direction"]), T64(f["lattice"]),
                                             torch.as_tensor(f["batch"]), int(f["batch"].max()) + 1, out_dims=(90, 1, 0, 3))
    for a, b in ((logits, f["logits"]), (vec, f["vec_out"]), (len0, f["len0"])):
        assert np.abs(a.numpy() - b).max() / np.abs(b).max() < 1e-11


def test_denoise_step_matches_live_reference(gold):
    s, w = gold("steps_c1_T1000.npz"), gold("weights_seed0.npz")
    with f64_default():
        tabs = R.DiffusionTables.build(1000, 90)
        for si, timestep in ((2, 500), (5, 1)):
            p = f"t{timestep}/"
            torch.manual_seed(2000 + si)
            G, N = s["num_atoms"].shape[0], s[p + "frac"].shape[0]
            z_len, z_frac, u = torch.randn(G, 3), torch.randn(N, 3), torch.rand(N, 90)
            out = R.denoise_step(_weights(w), tabs, T64(w["fourier_w"]), T64(s[p + "frac"]), torch.as_tensor(s[p + "types"]),
                                 T64(s[p + "lengths"]), T64(s["angles"]), torch.as_tensor(s["num_atoms"]), timestep,
                                 z_len, z_frac, u, 5.0, 8)
            d = np.abs(out[0].numpy() - s[p + "frac_next"])
            assert np.minimum(d, 1 - d).max() < 1e-11
            assert np.array_equal(out[1].numpy(), s[p + "types_next"])
            for a, key in ((out[2], "lengths_next"), (out[3], "lattice_next"), (out[4], "score"), (out[5], "logits"), (out[6], "len0")):
                assert np.abs(a.numpy() - s[p + key]).max() / np.abs(s[p + key]).max() < 1e-10, key


def test_reference_sample_first_and_last_step(gold):
    from arreau_b200.synthetic import calibrate_length_readout
    s, w = gold("sample_T11.npz"), gold("weights_seed0.npz")
    n_per, G = int(s["n_per"]), int(s["num_crystals"])
    # calibrate in fp64 exactly as oracle/gen_golden.py did (fp32-representable weights, fp64 products)
    sd = calibrate_length_readout({k: w[k].astype(np.float64) for k in w.files if k not in ("ori_grid", "fourier_w")}, n_per)
    with f64_default():
        W, tabs = _weights(w, sd=sd), R.DiffusionTables.build(11, 90)
        na = torch.full((G,), n_per)
        for k, timestep in ((0, 10), (9, 1)):
            o = R.denoise_step(W, tabs, T64(w["fourier_w"]), T64(s["step_frac"][k]), torch.as_tensor(s["step_types"][k]),
                               T64(s["step_lengths"][k]), T64(s["angles"]), na, timestep, T64(s["z_len"][k]),
                               T64(s["z_frac"][k]), T64(s["u_type"][k]), 5.0, 8)
            assert np.abs(o[4].numpy() - s["step_score"][k]).max() <= 1e-10 * max(1.0, np.abs(s["step_score"][k]).max())
            assert np.abs(o[5].numpy() - s["step_logits"][k]).max() <= 1e-10 * max(1.0, np.abs(s["step_logits"][k]).max())
        d = np.abs(o[0].numpy() - s["frac_x"])
        assert np.minimum(d, 1 - d).max() < 1e-11
        zt = np.array(list(range(1, 90)) + [2001])
        assert np.array_equal(zt[o[1].numpy()], s["atomic_numbers"])
        assert np.abs(o[3].numpy() - s["lattice"]).max() < 1e-10


def test_long_row_goldens_match_the_oracle(gold):
    """tests/golden/forward_longrows.npz (live reference, rows of up to 207 edges): the restatement reproduces the stored
    outputs and per-layer activations of the two C1-sized cases (the 2 x 200-atom case is pinned by gen_golden.py and
    exercised on the GPU; it takes too long for the CPU suite)."""
    f, w = gold("forward_longrows.npz"), gold("weights_seed0.npz")
    with f64_default():
        tabs = R.DiffusionTables.build(1000, 90)
        for case in ("c1_cap12", "c1_uncapped"):
            p = case + "/"
            na = torch.as_tensor(f[p + "num_atoms"])
            N = int(na.sum())
            t = torch.full((N,), int(f[p + "timestep"]))
            score, logits, len0, graph = R.predict_scores(
                _weights(w, float(f[p + "radius"])), tabs, T64(w["fourier_w"]), T64(f[p + "frac"]),
                torch.nn.functional.one_hot(torch.as_tensor(f[p + "types"]), 90), t, na, T64(f[p + "lengths"]),
                T64(f[p + "angles"]), float(f[p + "radius"]), int(f[p + "cap"]), return_graph=True)
            assert np.array_equal(graph[0][0].numpy(), f[p + "src"]) and np.array_equal(graph[0][1].numpy(), f[p + "dst"])
            assert np.bincount(f[p + "dst"]).max() > 8
            for a, key in ((score, "score"), (logits, "logits"), (len0, "len0")):
                assert np.abs(a.numpy() - f[p + key]).max() / np.abs(f[p + key]).max() < 1e-10, (case, key)
