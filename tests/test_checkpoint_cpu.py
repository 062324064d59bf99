"""Checkpoint round trip in the reference's on-disk format (CPU: parameter
containers only, no kernels)."""
from pathlib import Path

import torch
import yaml

from tests import helpers


def _write_reference_style_run(tmp_path, name, model_type):
    cls, kw, _ = helpers.MODEL_GOLDENS[name]
    _, sd = helpers.load_model_golden(name)
    run = tmp_path / 'run'
    (run / 'checkpoints').mkdir(parents=True)
    with open(run / 'model_kwargs.yaml', 'w') as f:
        yaml.dump(kw, f)
    with open(run / 'cmd_args.yaml', 'w') as f:
        yaml.dump({'model': model_type, 'learning_rate': 1e-3,
                   'weight_decay': 1e-4, 'use_1cycle': False,
                   'warm_restarts': False, 'egnn_attention': True,
                   'save_path': str(run), 'wandb_project': None,
                   'wandb_run': None}, f)
    # reference checkpoint dict keys: point_neural_network_base.py:509-516;
    # epochs are written in order, as training does (the newest file is the
    # one find_latest_checkpoint must pick: utils.py:33-45)
    import time
    torch.save({'model_state_dict': sd, 'optimiser_state_dict': {},
                'p_epoch': 1}, run / 'checkpoints' / 'pose_ckpt_epoch_1.pt')
    time.sleep(0.02)
    torch.save({'learning_rate': 1e-3, 'weight_decay': 1e-4, 'p_epoch': 3,
                'a_epoch': 0, 'model_state_dict': sd,
                'optimiser_state_dict': {}},
               run / 'checkpoints' / 'pose_ckpt_epoch_3.pt')
    return run, sd


def test_load_model_reads_reference_checkpoint_dir(tmp_path):
    from pointvs_b200.load_model import load_model
    run, sd = _write_reference_style_run(tmp_path, 'cfg3_k32', 'egnn')
    path, model, kwargs, cmd = load_model(run)
    assert path.name == 'pose_ckpt_epoch_3.pt'      # most recently written
    assert model.p_epoch == 3 and not model.training
    assert cmd['edge_attention'] is True
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    _, none_model, _, _ = load_model(run, fetch_args_only=True)
    assert none_model is None


def test_save_then_load_round_trip(tmp_path):
    import pointvs_b200 as pv
    kw = dict(dim_input=13, dim_output=1, k=16, num_layers=2, graphnorm=False,
              model_task='classification')
    torch.manual_seed(3)
    model = pv.MultitaskSatorrasEGNN(tmp_path / 'run2', 1e-3, 1e-4, None, None,
                                     **kw)
    assert (tmp_path / 'run2' / 'model_kwargs.yaml').is_file()
    model.p_epoch = 2
    model.save()
    ckpt = tmp_path / 'run2' / 'checkpoints' / 'pose_ckpt_epoch_2.pt'
    blob = torch.load(ckpt, weights_only=False)
    assert set(blob) == {'learning_rate', 'weight_decay', 'p_epoch', 'a_epoch',
                         'model_state_dict', 'optimiser_state_dict'}
    other = pv.MultitaskSatorrasEGNN(tmp_path / 'run2', 1e-3, 1e-4, None, None,
                                     silent=True, **kw)
    other.load_weights(ckpt)
    for (k, a), b in zip(model.state_dict().items(),
                         other.state_dict().values()):
        assert torch.equal(a, b), k


def test_legacy_attention_key_names_are_accepted(tmp_path):
    """Old checkpoints call att_mlp `edge_attention_mlp`
    (point_neural_network_base.py:520-526)."""
    from pointvs_b200.load_model import load_model
    run, sd = _write_reference_style_run(tmp_path, 'cfg3_k32', 'egnn')
    legacy = {k.replace('.att_mlp.', '.edge_attention_mlp.')
               .replace('.node_att_mlp.', '.node_attention_mlp.'): v
              for k, v in sd.items()}
    torch.save({'model_state_dict': legacy, 'optimiser_state_dict': {},
                'p_epoch': 9}, run / 'checkpoints' / 'pose_ckpt_epoch_9.pt')
    _, model, _, _ = load_model(run)
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_prediction_line_formats():
    import pointvs_b200 as pv
    kw = dict(dim_input=13, dim_output=1, k=16, num_layers=1, graphnorm=False)
    model = pv.SartorrasEGNN(Path('/tmp/pvs_test'), 0, 0, None, None,
                             silent=True, **kw)
    text = model._format_predictions(
        torch.tensor([0.25, 0.75]), torch.tensor([1.0, 0.0]),
        ['l0.parquet', 'l1.parquet'], ['r0.parquet', 'r1.parquet'])
    assert text == ('1.000 | 0.250 r0.parquet l0.parquet\n'
                    '0.000 | 0.750 r1.parquet l1.parquet\n')
    model.set_task('regression')
    text = model._format_predictions(torch.tensor([6.5]), None, ['l'], ['r'])
    assert text == '6.500 | r l\n'
