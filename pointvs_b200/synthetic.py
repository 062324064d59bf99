"""Deterministic synthetic protein-ligand complexes.

Atoms are uniform in a ball of density 0.065 per cubic Angstrom (what a real
pocket with hydrogens has: ~15.35 directed edges/atom at 4 A), sorted by
distance from the origin so the innermost `n_lig` atoms form the ligand.
Coordinates are rounded to float32 and stored as float64, so the fp64 graph
builder and the fp32 model see bit-identical values.  Features mirror
`--use_atomic_numbers --hydrogens --compact` (dim_input 13: one-hot type in
columns 0..11, column 12 = receptor flag; data_loaders.py:194-226 of the
reference).
"""
import numpy as np

DENSITY = 0.065
N_TYPES = 12
DIM_INPUT = 13


def synthetic_complex(seed, n_atoms=1000, n_lig=30, density=DENSITY):
    """Returns (coords float64 [N,3], bp int32 [N], feats float32 [N,13])."""
    rng = np.random.default_rng(seed)
    ball_r = (3.0 * n_atoms / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    direction = rng.normal(size=(n_atoms, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    radius = ball_r * rng.random(n_atoms) ** (1.0 / 3.0)
    coords = direction * radius[:, None]
    coords = coords[np.argsort(radius, kind='stable')]
    coords = coords.astype(np.float32).astype(np.float64)
    bp = np.ones(n_atoms, dtype=np.int32)
    bp[:n_lig] = 0
    types = rng.integers(0, N_TYPES, n_atoms)
    feats = np.zeros((n_atoms, DIM_INPUT), dtype=np.float32)
    feats[np.arange(n_atoms), types] = 1.0
    feats[:, DIM_INPUT - 1] = bp
    return coords, bp, feats


def synthetic_batch(first_seed, n_complexes, n_atoms=1000, n_lig=30,
                    ragged=False):
    """Packed batch: coords [sum N,3] f64, bp, feats, complex_ptr [B+1] int32.

    With `ragged`, complex i has n_atoms * U{0.8..1.2} atoms (seeded)."""
    coords, bps, feats, ptr = [], [], [], [0]
    for i in range(n_complexes):
        seed = first_seed + i
        n = n_atoms
        if ragged:
            n = int(np.random.default_rng(10_000_019 + seed).integers(
                int(0.8 * n_atoms), int(1.2 * n_atoms) + 1))
        c, b, f = synthetic_complex(seed, n, n_lig)
        coords.append(c)
        bps.append(b)
        feats.append(f)
        ptr.append(ptr[-1] + n)
    return (np.concatenate(coords), np.concatenate(bps),
            np.concatenate(feats), np.asarray(ptr, dtype=np.int32))


def synthetic_pocket_poses(first_pose, n_poses, n_pocket=800, n_lig=30,
                           density=DENSITY):
    """BASELINE configs[4] shape: ONE fixed pocket (seed 0, inner cavity
    removed) and `n_poses` ligand poses (seed = pose index) placed in the
    cavity.  Returns the same packed arrays as synthetic_batch."""
    rng = np.random.default_rng(0)
    n_all = n_pocket + n_lig
    ball_r = (3.0 * n_all / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    cavity_r = (3.0 * n_lig / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    direction = rng.normal(size=(n_pocket, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    u = rng.random(n_pocket)
    radius = (cavity_r ** 3 + u * (ball_r ** 3 - cavity_r ** 3)) ** (1.0 / 3.0)
    pocket = (direction * radius[:, None]).astype(np.float32).astype(np.float64)
    pocket_types = rng.integers(0, N_TYPES, n_pocket)
    coords, bps, feats, ptr = [], [], [], [0]
    for pose in range(first_pose, first_pose + n_poses):
        prng = np.random.default_rng(1_000_003 + pose)
        d = prng.normal(size=(n_lig, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        lig = d * (cavity_r * prng.random(n_lig) ** (1.0 / 3.0))[:, None]
        lig = lig.astype(np.float32).astype(np.float64)
        c = np.concatenate([lig, pocket])
        types = np.concatenate([prng.integers(0, N_TYPES, n_lig), pocket_types])
        bp = np.ones(n_all, dtype=np.int32)
        bp[:n_lig] = 0
        f = np.zeros((n_all, DIM_INPUT), dtype=np.float32)
        f[np.arange(n_all), types] = 1.0
        f[:, DIM_INPUT - 1] = bp
        coords.append(c), bps.append(bp), feats.append(f)
        ptr.append(ptr[-1] + n_all)
    return (np.concatenate(coords), np.concatenate(bps), np.concatenate(feats),
            np.asarray(ptr, dtype=np.int32))


POSE_BLOCK = 4096


def synthetic_pocket(n_pocket=800, n_lig=30, density=DENSITY):
    """The fixed pocket of `synthetic_pocket_poses` (seed 0): coords f64
    [n_pocket, 3], atom types int64 [n_pocket], and the cavity radius."""
    rng = np.random.default_rng(0)
    n_all = n_pocket + n_lig
    ball_r = (3.0 * n_all / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    cavity_r = (3.0 * n_lig / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    direction = rng.normal(size=(n_pocket, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    u = rng.random(n_pocket)
    radius = (cavity_r ** 3 + u * (ball_r ** 3 - cavity_r ** 3)) ** (1.0 / 3.0)
    pocket = (direction * radius[:, None]).astype(np.float32).astype(np.float64)
    return pocket, rng.integers(0, N_TYPES, n_pocket), cavity_r


def synthetic_ligand_poses(first_pose, n_poses, n_lig=30, density=DENSITY):
    """Ligand poses for a screening sweep over millions of poses: pose p is a
    pure function of p (block p // 4096 seeds one generator, the pose is row
    p % 4096 of its draw), generated a block at a time so that 1 M poses cost
    seconds, not a generator construction per pose.  Same distribution as the
    ligands of `synthetic_pocket_poses` (uniform in the pocket's cavity).
    Returns coords f64 [n_poses, n_lig, 3] (float32-representable) and atom
    types int16 [n_poses, n_lig]."""
    cavity_r = (3.0 * n_lig / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    coords = np.empty((n_poses, n_lig, 3), dtype=np.float64)
    types = np.empty((n_poses, n_lig), dtype=np.int16)
    done = 0
    while done < n_poses:
        p = first_pose + done
        blk, row = divmod(p, POSE_BLOCK)
        take = min(POSE_BLOCK - row, n_poses - done)
        rng = np.random.default_rng(2_000_003 + blk)
        d = rng.normal(size=(POSE_BLOCK, n_lig, 3))
        rad = cavity_r * rng.random((POSE_BLOCK, n_lig)) ** (1.0 / 3.0)
        ty = rng.integers(0, N_TYPES, (POSE_BLOCK, n_lig))
        d = d[row:row + take]
        d /= np.linalg.norm(d, axis=2, keepdims=True)
        c = d * rad[row:row + take, :, None]
        coords[done:done + take] = c.astype(np.float32).astype(np.float64)
        types[done:done + take] = ty[row:row + take]
        done += take
    return coords, types
