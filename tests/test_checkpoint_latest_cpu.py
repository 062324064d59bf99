"""find_latest_checkpoint follows the reference (utils.py:33-45): the most
recently created `<model_task>*.pt`, not the largest epoch number; and the
best-model metrics read the reference's predictions-file format."""
import os
import time

import pytest

from pointvs_b200 import base


def _touch(path, ctime_gap=0.02):
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_bytes(b'x')
    time.sleep(ctime_gap)


def test_latest_checkpoint_is_most_recently_created(tmp_path):
    ck = tmp_path / 'checkpoints'
    _touch(ck / 'pose_ckpt_epoch_1.pt')
    _touch(ck / 'pose_ckpt_epoch_2.pt')
    _touch(ck / 'affinity_ckpt_epoch_1.pt')     # written last, smaller epoch
    assert base.find_latest_checkpoint(tmp_path).name == 'affinity_ckpt_epoch_1.pt'
    assert base.find_latest_checkpoint(tmp_path, 'pose').name == 'pose_ckpt_epoch_2.pt'
    assert base.find_latest_checkpoint(tmp_path, 'affinity').name == \
        'affinity_ckpt_epoch_1.pt'
    with pytest.raises(RuntimeError):
        base.find_latest_checkpoint(tmp_path, 'both')
    with pytest.raises(FileNotFoundError):
        base.find_latest_checkpoint(tmp_path / 'nothing_here')


def test_top_n_and_pearson_on_predictions_file(tmp_path):
    f = tmp_path / 'pose_predictions.txt'
    f.write_text(
        '1 | 0.900 recA ligA_0\n0 | 0.200 recA ligA_1\n'     # top-1 hit
        '0 | 0.800 recB ligB_0\n1 | 0.300 recB ligB_1\n'     # top-1 miss
        '0 | 0.100 recC ligC_0\n0 | 0.050 recC ligC_1\n')    # no active
    assert base.top_n(f) == pytest.approx(1 / 3)
    assert base.top_n(f, n=2) == pytest.approx(2 / 3)
    g = tmp_path / 'affinity_predictions.txt'
    g.write_text(''.join(f'{y:.3f} | {y + 0.1 * ((-1) ** i):.3f} rec lig{i}\n'
                         for i, y in enumerate([4.0, 5.0, 6.0, 7.0, 8.0, 9.0])))
    r, p = base.get_regression_pearson(g)
    assert r > 0.99 and p < 0.05


def test_transform_names_leaves_node_attention_alone():
    d = {'layers.1.edge_attention_mlp.2.weight': 1,
         'layers.1.node_attention_mlp.0.weight': 2,
         'layers.2.node_att_mlp.2.bias': 3}
    out = base.PointNeuralNetworkBase._transform_names(d)
    assert list(out) == ['layers.1.att_mlp.0.weight',
                         'layers.1.node_att_mlp.0.weight',
                         'layers.2.node_att_mlp.2.bias']
