"""Mirror of diffusion/d3pm.py for the mask-absorbing chain: forward sampling (q_sample / get_xt), the reverse step
and the hybrid loss value."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from ..tables import build_tables


class D3PM(nn.Module):
    """diffusion/d3pm.py:8-72,198-215.  Only forward_type="mask" (diffusion_loss.py:77-82) is supported; the
    [T,Z,Z] matrices of the reference collapse to two numbers per timestep for this chain."""

    def __init__(self, x0_model=None, n_T: int = 1000, num_classes: int = 10, forward_type="mask",
                 hybrid_loss_coeff=0.001):
        super().__init__()
        if forward_type != "mask":
            raise NotImplementedError("only the mask-absorbing chain is on the denoising path")
        self.n_T, self.num_classses, self.eps = n_T, num_classes, 1e-6
        self.hybrid_loss_coeff = hybrid_loss_coeff
        self.tables = build_tables(n_T, num_classes)
        self.register_buffer("q_keep", self.tables.q_keep)
        self.register_buffer("q_to_mask", self.tables.q_to_mask)

    def q_sample(self, x_0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """d3pm.py:119-127: argmax(log(Qbar_t[x_0] + eps) + gumbel(noise)); noise is the [N,Z] uniform draw."""
        if not x_0.is_cuda:
            raise RuntimeError("arreau_b200 runs on CUDA tensors only (no CPU fallback)")
        dev = x_0.device
        x0 = x_0.to(torch.int64).contiguous()
        u = noise.to(dev, torch.float64).contiguous()
        tt = t.reshape(-1).to(dev, torch.int32).contiguous()
        q_keep, q_to_mask = self.q_keep.to(dev), self.q_to_mask.to(dev)
        out = torch.empty_like(x0)
        _lib.call("arreau_d3pm_q_sample", x0.data_ptr(), u.data_ptr(), tt.data_ptr(), q_keep.data_ptr(),
                  q_to_mask.data_ptr(), x0.shape[0], self.num_classses, out.data_ptr(),
                  torch.cuda.current_stream(dev).cuda_stream)
        return out

    def get_xt(self, x_0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """d3pm.py:139-143."""
        return self.q_sample(x_0, t, torch.rand((x_0.shape[0], self.num_classses), device=x_0.device))

    def reverse(self, x_t: torch.Tensor, predicted_x0_logits: torch.Tensor, t: torch.Tensor, noise=None):
        """argmax(q_posterior_logits + gumbel * (0.2 + 0.8 [t != 1])); `noise` injects torch.rand((N, Z))."""
        if not x_t.is_cuda:
            raise RuntimeError("arreau_b200 runs on CUDA tensors only (no CPU fallback)")
        dev = x_t.device
        types = x_t.to(torch.int64).contiguous()
        N, Z = types.shape[0], self.num_classses
        logits = predicted_x0_logits.to(dev, torch.float32).contiguous()
        u = torch.rand((N, Z), dtype=torch.float64, device=dev) if noise is None else noise.to(dev, torch.float64).contiguous()
        tt = t.reshape(-1).to(dev, torch.int32).contiguous()
        out = torch.empty_like(types)
        tb = self.tables
        q_keep, q_to_mask = self.q_keep.to(dev), self.q_to_mask.to(dev)   # keep the device copies alive over the call
        _lib.call("arreau_d3pm_reverse", types.data_ptr(), logits.data_ptr(), u.data_ptr(), tt.data_ptr(), 0,
                  q_keep.data_ptr(), q_to_mask.data_ptr(), tb.onestep_keep,
                  tb.onestep_to_mask, self.n_T, N, Z, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        return out
