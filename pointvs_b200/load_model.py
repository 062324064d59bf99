"""Rebuild a model from a reference-format checkpoint directory.

Mirror of /root/reference/point_vs/models/load_model.py:17-90: reads
`model_kwargs.yaml` and `cmd_args.yaml` next to `checkpoints/*.pt`, builds the
CUDA-backed class for `cmd_args['model']` ('egnn' or 'multitask') and loads the
state_dict unchanged (same parameter names and shapes as the reference).
"""
from pathlib import Path

import yaml

from .base import find_latest_checkpoint
from .egnn import MultitaskSatorrasEGNN, SartorrasEGNN

MODEL_CLASSES = {'egnn': SartorrasEGNN, 'multitask': MultitaskSatorrasEGNN}


def _load_yaml(path):
    with open(path, encoding='utf-8') as f:
        return yaml.safe_load(f) or {}


def load_model(model_path, silent=True, fetch_args_only=False, init_path=False,
               model_task=None):
    """Returns (checkpoint path, model or None, model_kwargs, cmd_line_args)."""
    model_path = Path(model_path).expanduser()
    if model_path.is_dir():
        model_path = find_latest_checkpoint(model_path, model_task=model_task)
    model_kwargs = _load_yaml(model_path.parents[1] / 'model_kwargs.yaml')
    cmd_line_args = _load_yaml(model_path.parents[1] / 'cmd_args.yaml')
    cmd_line_args.setdefault('node_attention', False)
    if 'edge_attention' not in cmd_line_args:
        cmd_line_args['edge_attention'] = cmd_line_args.get(
            'egnn_attention', False)
        model_kwargs['edge_attention'] = cmd_line_args['edge_attention']
    if fetch_args_only:
        return model_path, None, model_kwargs, cmd_line_args
    model_type = cmd_line_args['model']
    if model_type not in MODEL_CLASSES:
        raise NotImplementedError(
            f"model '{model_type}' is outside the B200 hot path "
            f"(supported: {sorted(MODEL_CLASSES)})")
    if init_path:
        save_path = Path(cmd_line_args['save_path'])
        project, run = cmd_line_args.get('wandb_project'), \
            cmd_line_args.get('wandb_run')
        if project is not None and run is not None:
            save_path = Path(save_path, project, run)
    else:
        save_path = Path()
    model = MODEL_CLASSES[model_type](
        save_path, learning_rate=cmd_line_args['learning_rate'],
        weight_decay=cmd_line_args['weight_decay'],
        use_1cycle=cmd_line_args.get('use_1cycle', False),
        warm_restarts=cmd_line_args.get('warm_restarts', False),
        regression_loss=cmd_line_args.get('regression_loss', 'mse'),
        silent=silent, **model_kwargs)
    model.load_weights(model_path, silent=silent)
    return model_path, model.eval(), model_kwargs, cmd_line_args
