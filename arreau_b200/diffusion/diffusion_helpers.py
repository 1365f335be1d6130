"""Mirror of diffusion/diffusion_helpers.py (hot-path part): same names, arguments and return values,
CUDA tensors in and out, every computation a kernel of libarreau_b200.so."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..tables import build_tables


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _topology(num_atoms: torch.Tensor, device):
    """atom_offset[G+1] i32 and crystal_of_atom[N] i32 from num_atoms[G] (device-side plumbing)."""
    na = num_atoms.to(device=device, dtype=torch.int64)
    off = torch.zeros(na.shape[0] + 1, dtype=torch.int64, device=device)
    torch.cumsum(na, 0, out=off[1:])
    coa = torch.repeat_interleave(torch.arange(na.shape[0], device=device), na)
    return off.to(torch.int32), coa.to(torch.int32)


def _cuda(t: torch.Tensor, dtype) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("arreau_b200 runs on CUDA tensors only (no CPU fallback)")
    return t.to(dtype).contiguous()


class GaussianFourierProjection(nn.Module):
    """diffusion/diffusion_helpers.py:14-25.  Only the frozen weight lives here; the projection itself is
    fused into the feature-assembly kernel (arreau_assemble_features)."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.gaussian_fourier_proj_w = nn.Parameter(torch.randn(embedding_size) * scale, requires_grad=False)


def frac_to_cart_coords(frac_coords: torch.Tensor, lattice: torch.Tensor, num_atoms: torch.Tensor) -> torch.Tensor:
    """diffusion/diffusion_helpers.py:223-230."""
    frac = _cuda(frac_coords, torch.float64)
    lat = _cuda(lattice, torch.float64)
    _, coa = _topology(num_atoms, frac.device)
    pos = torch.empty_like(frac)
    _lib.call("arreau_frac_to_cart", frac.data_ptr(), lat.data_ptr(), coa.data_ptr(), frac.shape[0], pos.data_ptr(),
              _stream(frac.device))
    return pos


def radius_graph_pbc(cart_coords, lattice, num_atoms, radius, max_num_neighbors_threshold, device=None,
                     topk_per_pair=None, remove_self_edges=True):
    """diffusion/diffusion_helpers.py:328-564.  Returns (edge_index[2,E] int64 (row 0 sender, row 1 receiver),
    cell_offsets[E,3], num_neighbors_image[G] int64, atomic_distance[E], neighbor_direction[E,3]); edges are
    receiver-major (i, j, cell).  Exact-distance ties under the cap are broken by ascending (j, cell)."""
    if topk_per_pair is not None:
        raise NotImplementedError("topk_per_pair is never used on the denoising path")
    pos = _cuda(cart_coords, torch.float64)
    dev = pos.device
    lat = _cuda(lattice, torch.float64)
    off, coa = _topology(num_atoms, dev)
    N, G = pos.shape[0], lat.shape[0]
    cap = int(max_num_neighbors_threshold)
    s = _stream(dev)
    raw = torch.empty(N, dtype=torch.int32, device=dev)
    deg = torch.empty(N, dtype=torch.int32, device=dev)
    row_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    nimg = torch.empty(G, dtype=torch.int64, device=dev)
    r2 = float(radius) * float(radius)
    _lib.call("arreau_graph_count", pos.data_ptr(), lat.data_ptr(), off.data_ptr(), coa.data_ptr(), N, G, r2, cap,
              int(bool(remove_self_edges)), raw.data_ptr(), deg.data_ptr(), nimg.data_ptr(), s)
    _lib.call("arreau_graph_scan", deg.data_ptr(), row_ptr.data_ptr(), N, s)
    E = int(row_ptr[N].item())          # the one host synchronisation: sizes the returned tensors
    src = torch.empty(E, dtype=torch.int32, device=dev)
    dst = torch.empty(E, dtype=torch.int32, device=dev)
    cell = torch.empty(E, dtype=torch.int8, device=dev)
    dist = torch.empty(E, dtype=torch.float64, device=dev)
    direction = torch.empty(E, 3, dtype=torch.float64, device=dev)
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    cell_offsets = torch.empty(E, 3, dtype=torch.float64, device=dev)
    _lib.call("arreau_graph_fill", pos.data_ptr(), lat.data_ptr(), off.data_ptr(), coa.data_ptr(), N, G, r2, cap,
              int(bool(remove_self_edges)), raw.data_ptr(), row_ptr.data_ptr(), E, src.data_ptr(), dst.data_ptr(),
              cell.data_ptr(), dist.data_ptr(), direction.data_ptr(), edge_index.data_ptr(), cell_offsets.data_ptr(),
              None, s)
    return edge_index, cell_offsets, nimg, dist, direction


class VE_pbc(nn.Module):
    """diffusion/diffusion_helpers.py:28-81: forward noising (training) and reverse step (sampling)."""

    def __init__(self, num_steps, sigma_min, sigma_max):
        super().__init__()
        self.T, self.sigma_min, self.sigma_max = num_steps, sigma_min, sigma_max
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            sig = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), self.T + 1))
        finally:
            torch.set_default_dtype(prev)
        self.register_buffer("sigmas", sig)

    def forward(self, frac_x0, t, lattice, num_atoms, noise=None, **kwargs):
        """helpers:43-63: returns (frac_noisy, wrapped_frac_eps_x, used_sigmas).  `noise` injects randn_like(frac_x0)."""
        frac0 = _cuda(frac_x0, torch.float64)
        dev = frac0.device
        eps = torch.randn_like(frac0) if noise is None else _cuda(noise, torch.float64)
        lat = _cuda(lattice.reshape(-1, 3, 3), torch.float64)
        tt = _cuda(t.reshape(-1), torch.int32)
        _, coa = _topology(num_atoms, dev)
        sigmas = self.sigmas.to(dev)
        noisy, target = torch.empty_like(frac0), torch.empty_like(frac0)
        _lib.call("arreau_ve_pbc_forward", frac0.data_ptr(), eps.data_ptr(), tt.data_ptr(), sigmas.data_ptr(),
                  lat.data_ptr(), coa.data_ptr(), frac0.shape[0], noisy.data_ptr(), target.data_ptr(), _stream(dev))
        return noisy, target, sigmas[tt.long()].view(-1, 1)

    def reverse(self, xt, epx_x, t, lattice=None, num_atoms=None, noise=None):
        """Returns (xt - eps (s_t^2 - s_{t-1}^2) + sqrt(s_{t-1}^2 (s_t^2 - s_{t-1}^2) / s_t^2) z) % 1.
        `noise` (optional) injects z; default torch.randn_like(xt) like the reference."""
        frac = _cuda(xt, torch.float64)
        dev = frac.device
        z = torch.randn_like(frac) if noise is None else _cuda(noise, torch.float64)
        score = _cuda(epx_x, torch.float32)
        tt = _cuda(t.reshape(-1), torch.int32)
        out = torch.empty_like(frac)
        sigmas = self.sigmas.to(dev)                  # keep the device copy alive over the call
        _lib.call("arreau_ve_pbc_reverse", frac.data_ptr(), score.data_ptr(), z.data_ptr(), tt.data_ptr(), 0,
                  sigmas.data_ptr(), frac.shape[0], out.data_ptr(), _stream(dev))
        return out


class VP_lattice(nn.Module):
    """diffusion/diffusion_helpers.py:134-199 (reverse_given_x0 only)."""

    def __init__(self, num_steps=1000, s=0.0001, power=2, clipmax=0.999):
        super().__init__()
        if (s, power, clipmax) != (0.0001, 2, 0.999):
            raise NotImplementedError("only the reference's schedule (diffusion_loss.py:30-36) is tabulated")
        self.tables = build_tables(num_steps, 2)
        self.register_buffer("alpha_bars", self.tables.vp_alpha_bars)
        self.register_buffer("betas", self.tables.vp_betas)
        self.register_buffer("sigmas", self.tables.vp_sigmas)

    def forward(self, h0, t, noise=None):
        """helpers:156-163: (sqrt(abar_t) h0 + sqrt(1 - abar_t) eps, eps); t is [G] or [G,1]."""
        lengths = _cuda(h0, torch.float64)
        dev = lengths.device
        eps = torch.randn_like(lengths) if noise is None else _cuda(noise, torch.float64)
        tt = _cuda(t.reshape(-1), torch.int32)
        ab = self.alpha_bars.to(dev, torch.float32)
        out = torch.empty_like(lengths)
        _lib.call("arreau_vp_lattice_forward", lengths.data_ptr(), eps.data_ptr(), tt.data_ptr(), ab.data_ptr(),
                  lengths.shape[0], out.data_ptr(), _stream(dev))
        return out, eps

    def reverse_given_x0(self, xt, pred_x0, t, noise=None):
        """`pred_x0` is the already scaled prediction (len0 * num_atoms, diffusion_loss.py:338)."""
        lengths = _cuda(xt, torch.float64)
        dev = lengths.device
        G = lengths.shape[0]
        ti = int(t.reshape(-1)[0])
        z = torch.randn_like(lengths) if noise is None else _cuda(noise, torch.float64)
        pred = _cuda(pred_x0, torch.float32)
        unit = torch.arange(G + 1, dtype=torch.int32, device=dev)   # num_atoms = 1: pred is already scaled
        tb = self.tables
        out = torch.empty_like(lengths)
        _lib.call("arreau_vp_lattice_reverse", lengths.data_ptr(), pred.data_ptr(), unit.data_ptr(), z.data_ptr(), ti,
                  float(tb.vp_cx0[ti]), float(tb.vp_cxt[ti]), float(tb.vp_denom[ti]), float(tb.vp_var[ti]), G,
                  out.data_ptr(), _stream(dev))
        return out
