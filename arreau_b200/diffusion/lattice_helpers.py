"""Mirror of diffusion/lattice_helpers.py (hot-path part)."""
from __future__ import annotations

import torch

from .. import _lib


def _dev(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("arreau_b200 runs on CUDA tensors only (no CPU fallback)")
    return t


def lattice_from_params(lengths: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """diffusion/lattice_helpers.py:55-105: (a,b,c), (alpha,beta,gamma as radians) -> [G,3,3] rows a,b,c."""
    lengths = _dev(lengths).to(torch.float64).contiguous()
    angles = angles.to(lengths.device, torch.float64).contiguous()
    G = lengths.shape[0]
    out = torch.empty(G, 3, 3, dtype=torch.float64, device=lengths.device)
    _lib.call("arreau_lattice_from_params", lengths.data_ptr(), angles.data_ptr(), G, out.data_ptr(),
              torch.cuda.current_stream(lengths.device).cuda_stream)
    return out


def matrix_to_params(matrix: torch.Tensor):
    """diffusion/lattice_helpers.py:16-35: (lengths[G,3], angles[G,3] in radians) of [G,3,3] row-vector lattices."""
    m = _dev(matrix).to(torch.float64).reshape(-1, 3, 3).contiguous()
    G = m.shape[0]
    lengths = torch.empty(G, 3, dtype=torch.float64, device=m.device)
    angles = torch.empty(G, 3, dtype=torch.float64, device=m.device)
    _lib.call("arreau_matrix_to_params", m.data_ptr(), G, lengths.data_ptr(), angles.data_ptr(),
              torch.cuda.current_stream(m.device).cuda_stream)
    return lengths, angles
