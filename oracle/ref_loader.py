"""Import the LIVE reference (unmodified files under /root/reference) with the stand-in
packages of oracle/shims on sys.path.  TEST INFRASTRUCTURE ONLY, and only usable in the
build container: /root/reference does not exist on the GPU box, so nothing under tests -m gpu,
smoke() or bench.py calls this."""
from __future__ import annotations

import argparse
import contextlib
import os
import sys

REFERENCE_ROOT = "/root/reference"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffusion"))


def import_reference():
    """Returns a namespace with the reference modules the hot path uses."""
    if not reference_available():
        raise RuntimeError("live reference not present (only in the build container)")
    for p in (SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    torch.set_default_dtype(torch.float64)  # main_diffusion_generate.py:27
    import diffusion.diffusion_helpers as helpers
    import diffusion.diffusion_loss as dloss
    import diffusion.d3pm as d3pm
    import diffusion.lattice_helpers as lattice_helpers
    import lightning_wrappers.diffusion as wrapper
    import ponita.models.ponita as ponita_model
    import ponita.nn.embedding as embedding
    import ponita.utils.windowing as windowing
    from diffusion.tools.atomic_number_table import AtomicNumberTable
    from diffusion.inference.visualize_crystal import VisualizationSetting
    from torch_geometric.data import Batch
    return argparse.Namespace(helpers=helpers, dloss=dloss, d3pm=d3pm, lattice_helpers=lattice_helpers,
                              wrapper=wrapper, ponita_model=ponita_model, embedding=embedding,
                              windowing=windowing, AtomicNumberTable=AtomicNumberTable,
                              VisualizationSetting=VisualizationSetting, Batch=Batch)


def default_args(T: int = 1000, radius: float = 5.0, max_neighbors: int = 8):
    """argparse defaults of main_diffusion.py:88-120 plus the Makefile:7 diffusion settings."""
    return argparse.Namespace(dataset="synthetic", lr=3e-4, weight_decay=0.0, epochs=1, warmup=0,
                              layer_scale=1e-6, train_augm=False, hidden_dim=128, layers=5, radius=radius,
                              num_ori=16, basis_dim=256, degree=3, widening_factor=4,
                              multiple_readouts=True, num_timesteps=T, max_neighbors=max_neighbors)


@contextlib.contextmanager
def stable_sort():
    """Force torch.sort(stable=True) while the reference runs: the
    reference's tie order under the default non-stable sort is implementation defined; the
    build's canonical rule is ascending (d2, j, cell)."""
    import torch
    orig = torch.sort

    def _sort(input, dim=-1, descending=False, stable=False, **kw):
        return orig(input, dim=dim, descending=descending, stable=True, **kw)

    torch.sort = _sort
    try:
        yield
    finally:
        torch.sort = orig
