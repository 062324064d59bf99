"""Row N2: the types-file / parquet loader (pointvs_b200/data.py) against what
the reference's own PygPointCloudDataset returned for the same files
(tests/golden/loader.npz, written by tests/golden/make_loader_golden.py).
Host side only: features, positions, labels, file lists bit-exact; the edge
list through the CPU oracle (the GPU tier repeats it through K1)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden.loader_configs import CONFIGS, ROOT

from oracle import radius_graph as rg
from pointvs_b200 import data


@pytest.fixture(scope='module')
def gold():
    return helpers.load_npz('loader.npz')


def _dataset(cfg, types='pose.types', task='classification', **extra):
    return data.ComplexDataset(ROOT, types_fname=ROOT / types,
                               model_task=task, rot=False, **CONFIGS[cfg],
                               **extra)


@pytest.mark.parametrize('cfg', sorted(CONFIGS))
def test_features_positions_and_edges_match_reference(cfg, gold):
    ds = _dataset(cfg)
    assert len(ds) == int(gold[f'{cfg}/n'])
    assert ds.feature_dim == int(gold[f'{cfg}/feature_dim'])
    for i in range(len(ds)):
        c = ds.load(i)
        np.testing.assert_array_equal(c.feats, gold[f'{cfg}/{i}/x'])
        np.testing.assert_array_equal(c.coords.astype(np.float32),
                                      gold[f'{cfg}/{i}/pos'])
        assert str(ds.ligand_fnames[i]) == str(gold[f'{cfg}/{i}/lig'])
        assert str(ds.receptor_fnames[i]) == str(gold[f'{cfg}/{i}/rec'])
        assert int(ds.label(i)) == int(gold[f'{cfg}/{i}/y'])
        _, row, col, attr = rg.radius_graph(c.coords, c.bp, ds.inter_radius,
                                            ds.intra_radius)
        ei = gold[f'{cfg}/{i}/edge_index']
        np.testing.assert_array_equal(row, ei[0])
        np.testing.assert_array_equal(col, ei[1])
        np.testing.assert_array_equal(attr, gold[f'{cfg}/{i}/edge_attr'])


def test_regression_labels_and_missing_files(gold):
    for task in ('regression', 'multi_regression'):
        ds = _dataset('smina_r10_e4', 'affinity.types', task)
        assert len(ds) == int(gold[f'{task}/n'])      # missing ligand dropped
        for i in range(len(ds)):
            y = np.asarray(ds.label(i), dtype=np.float32)
            np.testing.assert_array_equal(y, gold[f'{task}/{i}/y'])
            assert str(ds.ligand_fnames[i]) == str(gold[f'{task}/{i}/lig'])
        assert ds.sampler is None


def test_sampler_weights_match_reference(gold):
    ds = _dataset('smina_r10_e4')
    np.testing.assert_array_equal(np.asarray(ds.labels), gold['labels'])
    np.testing.assert_array_equal(ds.sample_weights.numpy(),
                                  gold['sample_weights'])
    assert isinstance(ds.sampler, torch.utils.data.WeightedRandomSampler)


def test_two_column_types_file_has_no_labels():
    ds = _dataset('smina_r10_e4', 'nolabel.types')
    assert len(ds) == 3 and ds.label(0) is None and ds.sampler is None


def test_types_parser_edge_cases(tmp_path):
    f = tmp_path / 'a.types'
    f.write_text('1 0.25 3.5 r/a.parquet l/b.parquet # c\n'
                 '\n'
                 '0 7 r/c.parquet l/d.parquet 12.5 1.0\n'
                 'r/e.parquet l/f.parquet\n')
    labels, rmsds, recs, ligs = data.parse_classification_types(f)
    assert labels == [1, 0, None]
    assert rmsds == [3.5, 7.0, None]
    assert recs == ['r/a.parquet', 'r/c.parquet', 'r/e.parquet']
    # the reference keeps the LAST non-numeric token as the ligand
    assert ligs == ['c', 'l/d.parquet', 'l/f.parquet']


def test_label_by_rmsd(tmp_path):
    lines = (ROOT / 'pose.types').read_text().splitlines()
    ds = _dataset('smina_r10_e4', max_active_rms_distance=1.5,
                  min_inactive_rms_distance=2.0)
    # rmsds in pose.types: 0.0 -1 1.0 -1 2.0 -1 3.0 ; negatives are dropped,
    # < 1.5 active, >= 2.0 inactive
    assert [ln for ln in lines if ln.strip()] and len(ds) == 4
    assert list(ds.labels) == [1, 1, 0, 0]


def test_unsupported_options_fail_loudly():
    for kw in (dict(prune=True), dict(bp=0), dict(p_noise=0.1),
               dict(augmented_active_count=2), dict(p_remove_entity=0.5),
               dict(include_strain_info=True)):
        with pytest.raises(NotImplementedError):
            _dataset('smina_r10_e4', **kw)
    with pytest.raises(NotImplementedError):      # as the reference
        data.ComplexDataset(ROOT, types_fname=ROOT / 'pose.types',
                            polar_hydrogens=True, use_atomic_numbers=False)
    with pytest.raises(FileNotFoundError):
        data.ComplexDataset(ROOT / 'nope', types_fname=ROOT / 'pose.types')


def test_atomic_number_table_matches_reference_layout():
    table, n = data.atomic_number_table(True)
    assert n == 12 and table[1] == 11 and table[6] == 0 and table[17] == 6
    assert table[35] == table[53] == 7 and table[30] == 10
    table, n = data.atomic_number_table(False)
    assert n == 11 and 1 not in table


def test_bit_vector_unmapped_element_quirk():
    # an element outside the table gets index n_features: in compact mode
    # that wraps to column 0 with the molecule bit incremented (reference
    # arithmetic, preprocessing.py:231-234)
    v = data.make_bit_vector(np.array([11, 22, 3]), 11, compact=True)
    assert v.shape == (3, 12)
    assert v[0, 0] == 1 and v[0, -1] == 1
    assert v[1, 0] == 1 and v[1, -1] == 2
    assert v[2, 3] == 1 and v[2, -1] == 0
    with pytest.raises(ValueError):
        data.make_bit_vector(np.array([22]), 11, compact=False)


def test_loader_batching_order_and_len():
    ds = _dataset('smina_r10_e4')
    dl = data.PackedLoader(ds, batch_size=3, num_workers=0)
    assert len(dl) == 3
    assert dl._batches() == [[0, 1, 2], [3, 4, 5], [6]]
    assert Path(ds.ligand_fnames[6]).name == 'lig_0.parquet'


def test_random_rotation_is_a_rotation():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(50, 3))
    y = data.uniform_random_rotation(x, rng)
    d0 = np.linalg.norm(x[:, None] - x[None], axis=-1)
    d1 = np.linalg.norm(y[:, None] - y[None], axis=-1)
    np.testing.assert_allclose(d0, d1, atol=1e-12)


def test_loader_worker_processes_match_threads():
    """processes=True (spawned loader workers, as the reference's DataLoader
    workers) prepares exactly what the in-process path prepares, in order."""
    for device_crop in (False, True):
        ds = _dataset('atomic_h_r6_e3', device_crop=device_crop)
        ref = list(data.PackedLoader(ds, batch_size=3, num_workers=0).prepared())
        dl = data.PackedLoader(ds, batch_size=3, num_workers=2, processes=True)
        try:
            got = list(dl.prepared())
        finally:
            dl.close()
        assert [i for i, _ in got] == [i for i, _ in ref]
        for (_, a), (_, b) in zip(got, ref):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                np.testing.assert_array_equal(x.coords, y.coords)
                if device_crop:
                    np.testing.assert_array_equal(x.emit, y.emit)
                    np.testing.assert_array_equal(x.code, y.code)
                else:
                    np.testing.assert_array_equal(x.feats, y.feats)
                    np.testing.assert_array_equal(x.bp, y.bp)
