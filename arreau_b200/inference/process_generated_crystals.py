"""Mirror of diffusion/inference/process_generated_crystals.py:8-33: the wire format of generated crystals.

Layout (group "crystals"): frac_x[N,3] f64, atomic_numbers[N], lattice[G,3,3] f64, idx_start[G], num_atoms[G].
HDF5 through h5py when it is installed (the reference's format, byte-compatible readers); otherwise -- h5py is not
part of this image -- the same five arrays under the same keys ("crystals/<name>") in a NumPy .npz archive."""
from __future__ import annotations

import os

import numpy as np

from ..diffusion.diffusion_loss import SampleResult

KEYS = ("frac_x", "atomic_numbers", "lattice", "idx_start", "num_atoms")


def _have_h5py() -> bool:
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


def save_sample_results_to_hdf5(crystals: SampleResult, filename: str) -> str:
    """process_generated_crystals.py:8-15.  Returns the path actually written (".npz" replaces ".h5" without h5py)."""
    os.makedirs(os.path.dirname(os.path.abspath(filename)) or ".", exist_ok=True)
    if _have_h5py() and not filename.endswith(".npz"):
        import h5py
        with h5py.File(filename, "w") as file:
            group = file.create_group("crystals")
            for k in KEYS:
                group.create_dataset(k, data=getattr(crystals, k))
        return filename
    path = filename if filename.endswith(".npz") else os.path.splitext(filename)[0] + ".npz"
    np.savez(path, **{f"crystals/{k}": np.asarray(getattr(crystals, k)) for k in KEYS})
    return path


def load_sample_results_from_hdf5(filename: str) -> SampleResult:
    """process_generated_crystals.py:18-31 (path taken as given)."""
    if filename.endswith(".npz") or not _have_h5py():
        path = filename if filename.endswith(".npz") else os.path.splitext(filename)[0] + ".npz"
        z = np.load(path)
        return SampleResult(**{k: z[f"crystals/{k}"] for k in KEYS})
    import h5py
    with h5py.File(filename, "r") as file:
        return SampleResult(**{k: file["crystals"][k][:] for k in KEYS})


def get_crystal_indexes(sample_result: SampleResult, sample_idx: int):
    """process_generated_crystals.py:34-38."""
    start = sample_result.idx_start[sample_idx]
    return start, start + sample_result.num_atoms[sample_idx]
