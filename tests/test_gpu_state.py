"""K8/K9 parity: lattice, frac->cart, VE / VP / D3PM reverse steps against the golden KATs and the oracle."""
import numpy as np
import pytest
import torch

from conftest import f64_default, rel_err

pytestmark = pytest.mark.gpu


def test_lattice_and_frac_to_cart_kat(device, gold):
    from arreau_b200.diffusion.diffusion_helpers import frac_to_cart_coords
    from arreau_b200.diffusion.lattice_helpers import lattice_from_params
    k = gold("kat.npz")
    lat = lattice_from_params(torch.as_tensor(k["lat_lengths"], device=device), torch.as_tensor(k["lat_angles"], device=device))
    # device libm sin/cos/acos may differ from the host's in the last ulp
    assert np.allclose(lat.cpu().numpy(), k["lat_matrix"], rtol=1e-14, atol=1e-15)
    cart = frac_to_cart_coords(torch.as_tensor(k["f2c_frac"], device=device),
                               torch.as_tensor(k["lat_matrix"], device=device), torch.tensor([1, 1, 1], device=device))
    assert np.array_equal(cart.cpu().numpy(), k["f2c_cart"])       # same fp64 operation order: bit exact


@pytest.mark.parametrize("timestep", [999, 500, 2, 1])
def test_update_kernels_against_oracle(device, timestep):
    from arreau_b200.diffusion.d3pm import D3PM
    from arreau_b200.diffusion.diffusion_helpers import VE_pbc, VP_lattice
    from oracle import restatement as R
    T, Z, N, G = 1000, 90, 257, 19
    g = torch.Generator().manual_seed(timestep)
    with f64_default():
        tabs = R.DiffusionTables.build(T, Z)
    frac = torch.rand(N, 3, generator=g, dtype=torch.float64) * 3 - 1
    score = torch.randn(N, 3, generator=g).float()
    z = torch.randn(N, 3, generator=g, dtype=torch.float64)
    t = torch.full((N,), timestep)
    ref = R.ve_pbc_reverse(tabs, frac, score.double(), t, z)
    got = VE_pbc(T, 0.001, 1.0).to(device).reverse(frac.to(device), score.to(device), t.to(device), None, None, noise=z.to(device))
    d = np.abs(got.cpu().numpy() - ref.numpy())
    assert np.minimum(d, 1 - d).max() < 1e-15

    lengths = torch.randn(G, 3, generator=g, dtype=torch.float64) * 2 + 6
    pred = (torch.randn(G, 3, generator=g) * 2 + 6).float()
    zl = torch.randn(G, 3, generator=g, dtype=torch.float64)
    tv = torch.tensor([timestep])
    ref = R.vp_lattice_reverse_given_x0(tabs, lengths, pred.double(), tv, zl)
    got = VP_lattice(T).reverse_given_x0(lengths.to(device), pred.to(device), tv.to(device), noise=zl.to(device))
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-15

    types = torch.randint(0, Z, (N,), generator=g)
    types[::3] = Z - 1
    logits = (torch.randn(N, Z, generator=g) * 3).float()
    u = torch.rand(N, Z, generator=g, dtype=torch.float64)
    ref = R.d3pm_reverse(tabs, types, logits.double(), t, u)
    got = D3PM(None, T, Z, "mask").reverse(types.to(device), logits.to(device), t.to(device), noise=u.to(device))
    assert np.array_equal(got.cpu().numpy(), ref.numpy())


def test_philox_noise_statistics(device, packed_weights):
    from arreau_b200 import _lib
    import ctypes as C
    G, N, Z = 64, 4096, 90
    zl = torch.empty(G, 3, dtype=torch.float64, device=device)
    zf = torch.empty(N, 3, dtype=torch.float64, device=device)
    u = torch.empty(N, Z, dtype=torch.float64, device=device)
    _lib.call("arreau_step_noise", C.c_uint64(7), 3, G, N, Z, zl.data_ptr(), zf.data_ptr(), u.data_ptr(), 0)
    torch.cuda.synchronize()
    assert abs(zf.mean().item()) < 0.03 and abs(zf.std().item() - 1) < 0.03
    assert 0 <= u.min().item() and u.max().item() < 1 and abs(u.mean().item() - 0.5) < 0.01
    u2 = torch.empty_like(u)
    _lib.call("arreau_step_noise", C.c_uint64(7), 4, G, N, Z, zl.data_ptr(), zf.data_ptr(), u2.data_ptr(), 0)
    torch.cuda.synchronize()
    assert not torch.equal(u, u2)
