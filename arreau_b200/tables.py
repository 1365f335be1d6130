"""Init-time tables of the diffusion processes (host side, torch CPU).

They are built with the very torch expressions the reference uses at construction time so the values
are the reference's bit for bit, including its mixed fp32/fp64 VP schedule:
  VE sigmas      diffusion/diffusion_helpers.py:38-41
  VP schedule    diffusion/diffusion_helpers.py:141-154
  D3PM matrices  diffusion/d3pm.py:25-59 (forward_type="mask", diffusion_loss.py:77-82)
Constants: diffusion/diffusion_loss.py:30-36.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

POS_SIGMA_MIN = 0.001
POS_SIGMA_MAX = 1.0
LATTICE_POWER = 2
LATTICE_CLIPMAX = 0.999
MASK_RATE = 0.02   # diffusion/d3pm.py:34
D3PM_EPS = 1e-6    # diffusion/d3pm.py:23


@dataclass
class DiffusionTables:
    T: int
    Z: int
    ve_sigmas: torch.Tensor        # [T+1] f64
    vp_alpha_bars: torch.Tensor    # [T+1] f32 (stays fp32 in the reference)
    vp_betas: torch.Tensor         # [T+1] f64
    vp_sigmas: torch.Tensor        # [T+1] f64
    # coefficients of VP_lattice.reverse_given_x0 per timestep (helpers:185-199), fp64 after promotion
    vp_cx0: torch.Tensor
    vp_cxt: torch.Tensor
    vp_denom: torch.Tensor
    vp_var: torch.Tensor
    # mask chain: Qbar_t[0,0] and Qbar_t[0,Z-1], index t-1
    q_keep: torch.Tensor           # [T] f64
    q_to_mask: torch.Tensor        # [T] f64
    onestep_keep: float
    onestep_to_mask: float


def _in_f64(fn):
    def wrapped(*a, **k):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)   # main_diffusion_generate.py:27
        try:
            return fn(*a, **k)
        finally:
            torch.set_default_dtype(prev)
    return wrapped


@_in_f64
def build_tables(T: int, Z: int) -> DiffusionTables:
    ve_sigmas = torch.exp(torch.linspace(np.log(POS_SIGMA_MIN), np.log(POS_SIGMA_MAX), T + 1))
    t = torch.arange(0, T + 1, dtype=torch.float)
    s = 0.0001
    f_t = torch.cos((np.pi / 2) * ((t / T) + s) / (1 + s)) ** LATTICE_POWER
    alpha_bars = f_t / f_t[0]
    betas = torch.cat([torch.zeros([1]), 1 - (alpha_bars[1:] / alpha_bars[:-1])], dim=0)
    betas = betas.clamp_max(LATTICE_CLIPMAX)
    sigmas = torch.sqrt(betas[1:] * ((1 - alpha_bars[:-1]) / (1 - alpha_bars[1:])))
    sigmas = torch.cat([torch.zeros([1]), sigmas], dim=0)
    # reverse_given_x0 coefficients for t = 1..T (index 0 unused)
    idx = torch.arange(1, T + 1)
    denom = 1 - alpha_bars[idx]
    alpha_t = 1 - betas[idx]
    cx0 = torch.sqrt(alpha_bars[idx - 1]) * betas[idx]
    cxt = torch.sqrt(alpha_t) * (1 - alpha_bars[idx - 1])
    var = (1 - alpha_bars[idx - 1]) * betas[idx] / denom
    pad = lambda v: torch.cat([torch.zeros(1, dtype=torch.float64), v.to(torch.float64)])  # noqa: E731
    # D3PM mask chain: same matrix products as the reference, only row 0 is kept (every non-mask row
    # has the same diagonal / mask-column entries; the mask row is (0, .., 0, 1)).
    mat = torch.zeros(Z, Z)
    mat[:, -1] = torch.full((Z,), MASK_RATE)
    mat.diagonal().fill_(1 - MASK_RATE)
    mat[-1, -1] = 1
    q = mat.clone()
    keep, to_mask = [q[0, 0].item()], [q[0, Z - 1].item()]
    for _ in range(1, T):
        q = q @ mat
        keep.append(q[0, 0].item())
        to_mask.append(q[0, Z - 1].item())
    return DiffusionTables(T=T, Z=Z, ve_sigmas=ve_sigmas, vp_alpha_bars=alpha_bars, vp_betas=betas, vp_sigmas=sigmas,
                           vp_cx0=pad(cx0), vp_cxt=pad(cxt), vp_denom=pad(denom), vp_var=pad(var),
                           q_keep=torch.tensor(keep), q_to_mask=torch.tensor(to_mask),
                           onestep_keep=float(mat[0, 0]), onestep_to_mask=float(mat[0, Z - 1]))
