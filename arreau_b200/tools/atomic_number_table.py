"""Mirror of diffusion/tools/atomic_number_table.py (the part the sampling / checkpoint path touches)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

# element symbols by atomic number (the reference uses pymatgen's Element(symbol).number, tools/atomic_number_table.py:88)
_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
            "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
            "Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr").split()
ATOMIC_NUMBER = {s: i + 1 for i, s in enumerate(_SYMBOLS)}


class AtomicNumberTable:
    """tools/atomic_number_table.py:7-26."""
    MASK_ATOMIC_NUMBER = 2001

    def __init__(self, zs: Sequence[int]):
        self.zs = list(zs)

    def __len__(self) -> int:
        return len(self.zs)

    def __str__(self):
        return f"AtomicNumberTable: {tuple(s for s in self.zs)}"

    def index_to_z(self, index: int) -> int:
        return self.zs[index]

    def z_to_index(self, atomic_number: int) -> int:
        return self.zs.index(atomic_number)


def atomic_numbers_to_indices(z_table: AtomicNumberTable, atomic_numbers) -> torch.Tensor:
    """tools/atomic_number_table.py:37-43."""
    return torch.tensor([z_table.z_to_index(int(z)) for z in np.asarray(atomic_numbers).reshape(-1)], dtype=torch.long)


def atomic_symbols_to_indices(z_table: AtomicNumberTable, atomic_symbols) -> torch.Tensor:
    """tools/atomic_number_table.py:84-89."""
    return atomic_numbers_to_indices(z_table, [ATOMIC_NUMBER[s] for s in atomic_symbols])
