"""Configurations shared by make_loader_golden.py and the loader tests."""
from pathlib import Path

ROOT = Path(__file__).resolve().parent / 'complexes'

CONFIGS = {
    # the reference CLI defaults for the egnn family: sminatypes, no hydrogens
    'smina_r10_e4': dict(radius=10, edge_radius=4, polar_hydrogens=False,
                         use_atomic_numbers=False, compact=True),
    'atomic_h_r6_e3': dict(radius=6, edge_radius=3, polar_hydrogens=True,
                           use_atomic_numbers=True, compact=True),
    'atomic_noh_r8_e4': dict(radius=8, edge_radius=4, polar_hydrogens=False,
                             use_atomic_numbers=True, compact=True),
    'smina_wide_r7_bonds': dict(radius=7, edge_radius=5, polar_hydrogens=False,
                                use_atomic_numbers=False, compact=False,
                                estimate_bonds=True),
    'smina_ext_r5_e0': dict(radius=5, edge_radius=0, polar_hydrogens=False,
                            use_atomic_numbers=False, compact=True,
                            extended_atom_types=True),
}
