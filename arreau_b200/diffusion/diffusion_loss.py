"""Mirror of diffusion/diffusion_loss.py: DiffusionLoss.predict_scores / sample with the reference's
signatures.  The sampler keeps the whole state on the GPU and runs each step as one C call
(arreau_denoise_step); the reference's torch CPU RNG stream (angles via numpy, randn lengths, randn frac, then
per step randn_like(lengths), randn_like(frac), rand((N,Z)) -- SURVEY 3.1) is drawn on the host exactly as the
reference draws it unless `device_noise=True` asks for the in-kernel Philox generator."""
from __future__ import annotations

from typing import Optional

import attr
import numpy as np
import torch
import torch.nn.functional as F

from ..engine import DenoiseEngine
from ..tables import POS_SIGMA_MAX, POS_SIGMA_MIN, build_tables
from .d3pm import D3PM
from .diffusion_helpers import VE_pbc, VP_lattice

pos_sigma_min = POS_SIGMA_MIN
pos_sigma_max = POS_SIGMA_MAX


@attr.s(auto_attribs=True, frozen=True)
class SampleResult:
    """diffusion/diffusion_loss.py:39-49."""
    frac_x: np.ndarray = None
    atomic_numbers: np.ndarray = None
    lattice: np.ndarray = None
    idx_start: np.ndarray = None
    num_atoms: np.ndarray = None


def sample_bravais_angles(lattice_type: str = "monoclinic"):
    """diffusion/diffusion_helpers.py:739-774, monoclinic branch (:752-755): degrees, consumed as radians (B5)."""
    if lattice_type != "monoclinic":
        raise NotImplementedError(lattice_type)
    return [90.0, np.random.uniform(90, 180), 90.0]


class DiffusionLoss(torch.nn.Module):
    """diffusion/diffusion_loss.py:67-93.  `args` needs .radius, .max_neighbors, .num_timesteps."""

    def __init__(self, args, num_atomic_states: int, precision: str = "fp32"):
        super().__init__()
        self.cutoff, self.max_neighbors, self.T = args.radius, args.max_neighbors, args.num_timesteps
        self.num_atomic_states = num_atomic_states
        self.precision = precision
        self.tables = build_tables(self.T, num_atomic_states)
        self.pos_diffusion = VE_pbc(self.T, sigma_min=pos_sigma_min, sigma_max=pos_sigma_max)
        self.d3pm = D3PM(x0_model=None, n_T=self.T, num_classes=num_atomic_states, forward_type="mask")
        self.lattice_diffusion = VP_lattice(num_steps=self.T)
        self._engine = None
        self._engine_key = None

    # -- engine cache: one per (model weights, topology) --------------------------------------------
    def engine_for(self, model, t_emb_weights, num_atoms, device, debug=False) -> DenoiseEngine:
        net = getattr(model, "model", model)          # PONITA_DIFFUSION.model or a bare PonitaFiberBundle
        na = tuple(int(v) for v in torch.as_tensor(num_atoms).reshape(-1).tolist())
        key = (id(net), na, str(device), debug, self.precision)
        if self._engine is None or self._engine_key != key:
            packed = net._packed if getattr(net, "_packed", None) is not None and net._packed.device == torch.device(device) \
                else net.pack(device)
            fw = t_emb_weights.gaussian_fourier_proj_w if hasattr(t_emb_weights, "gaussian_fourier_proj_w") else t_emb_weights
            self._engine = DenoiseEngine(packed, self.tables, fw, na, self.cutoff, self.max_neighbors,
                                         precision=self.precision, debug=debug, device=device)
            self._engine_key = key
        return self._engine

    def predict_scores(self, noisy_frac_x, noisy_atom_types, t_feat, num_atoms, noisy_lengths, angles, model, batch,
                       t_emb_weights):
        """diffusion/diffusion_loss.py:112-197.  noisy_atom_types is the one-hot [N,Z] of the reference."""
        dev = noisy_frac_x.device
        if dev.type != "cuda":
            raise RuntimeError("arreau_b200 runs on CUDA tensors only (no CPU fallback)")
        eng = self.engine_for(model, t_emb_weights, num_atoms, dev)
        types = noisy_atom_types.argmax(-1) if noisy_atom_types.dim() == 2 else noisy_atom_types
        eng.set_state(noisy_frac_x, types, noisy_lengths, angles)
        tt = t_feat.reshape(-1)
        score, logits, len0 = eng.predict_scores(tt)
        dt = noisy_frac_x.dtype
        return score.to(dt).clone(), logits.to(dt).clone(), len0.to(dt).clone()

    @torch.no_grad()
    def sample(self, *, model, z_table, t_emb_weights, num_atoms_per_sample: int, num_samples_in_batch: int,
               vis_name: str = "", visualization_setting=None, show_bonds: bool = False,
               constant_atoms: Optional[torch.Tensor] = None, num_atoms: Optional[torch.Tensor] = None,
               device="cuda", device_noise: bool = False, seed: int = 0, step_callback=None) -> SampleResult:
        """diffusion/diffusion_loss.py:277-377.  Extra keyword-only options (defaults reproduce the reference):
        `num_atoms` a per-crystal atom-count vector, `device_noise` Philox noise generated on the GPU,
        `step_callback(timestep, engine)` called after each step (e.g. to log E/N)."""
        Z = len(z_table)
        G = num_samples_in_batch
        dd = torch.float64                                     # main_diffusion_generate.py:27
        angles = torch.tensor([sample_bravais_angles("monoclinic") for _ in range(G)], dtype=dd)
        lengths = torch.randn([G, 3], dtype=dd)
        na = torch.full((G,), num_atoms_per_sample) if num_atoms is None else torch.as_tensor(num_atoms).reshape(-1)
        N = int(na.sum())
        frac_x = torch.randn([N, 3], dtype=dd) * pos_sigma_max
        atom_types = constant_atoms if constant_atoms is not None else torch.full((N,), Z - 1)
        eng = self.engine_for(model, t_emb_weights, na, torch.device(device))
        eng.set_state(frac_x, atom_types, lengths, angles)
        for step_idx, timestep in enumerate(reversed(range(1, self.T))):
            if device_noise:
                eng.draw_noise(seed, step_idx)
            else:   # the reference's draw order per step: lengths, frac, types
                eng.set_noise(torch.randn(G, 3, dtype=dd), torch.randn(N, 3, dtype=dd), torch.rand(N, Z, dtype=dd))
            eng.step(timestep, update_types=constant_atoms is None)
            if step_callback is not None:
                step_callback(timestep, eng)
        types = eng.types.cpu()
        zs = np.asarray(z_table.zs if hasattr(z_table, "zs") else z_table)
        return SampleResult(num_atoms=na.numpy(), frac_x=eng.frac.cpu().numpy(), atomic_numbers=zs[types.numpy()],
                            lattice=eng.lattice.cpu().numpy())
