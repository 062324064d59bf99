"""Scoring entry point, mirror of /root/reference/point_vs/inference.py.

    python -m pointvs_b200.inference <model_checkpoint> <test_types> \
        <test_data_root> [--model_task pose|affinity]

Loads a reference-format checkpoint into the CUDA-backed model, scores the
types file and writes `predictions_<types>-<ckpt>.txt` next to the checkpoint
in the reference's line format.  The parquet/types-file data loader is the
reference's own (`point_vs.preprocessing.data_loaders`, outside the hot path):
it must be importable, or a loader factory can be passed programmatically.
"""
import argparse
from pathlib import Path

from .load_model import load_model


def _reference_loader_factory():
    try:
        from point_vs.preprocessing.data_loaders import (  # noqa: PLC0415
            get_data_loader, PygPointCloudDataset)
    except ImportError as exc:   # pragma: no cover - depends on environment
        raise ImportError(
            'pointvs_b200.inference uses PointVS\'s own parquet/types data '
            'loader (point_vs.preprocessing.data_loaders); install PointVS or '
            'pass loader_factory=...') from exc

    def factory(data_root, **kwargs):
        return get_data_loader(data_root, dataset_class=PygPointCloudDataset,
                               **kwargs)
    return factory


def get_model_and_test_dl(checkpoint_path, test_types, test_data_root,
                          model_task=None, loader_factory=None):
    """Same contract as inference.py:35-74 of the reference."""
    checkpoint_path, model, model_kwargs, cmd_line_args = load_model(
        checkpoint_path, silent=False, model_task=model_task)
    if model_task is None:
        model_task_ = cmd_line_args.get('model_task', 'classification')
    else:
        model_task_ = {'pose': 'classification',
                       'affinity': 'regression'}[model_task]
        is_multi = cmd_line_args.get(
            'multimulti_target_affinity',
            cmd_line_args.get('model_task', 'multi_regression'))
        if is_multi and model_task_ == 'regression':
            model_task_ = 'multi_regression'
    model.set_task(model_task_)
    factory = loader_factory or _reference_loader_factory()
    test_dl = factory(
        test_data_root, receptors=None, compact=cmd_line_args['compact'],
        use_atomic_numbers=cmd_line_args['use_atomic_numbers'],
        radius=cmd_line_args['radius'],
        polar_hydrogens=cmd_line_args['hydrogens'],
        batch_size=cmd_line_args['batch_size'], types_fname=test_types,
        edge_radius=cmd_line_args['edge_radius'],
        estimate_bonds=cmd_line_args.get('estimate_bonds', False),
        prune=cmd_line_args.get('prune', False), rot=False, mode='val',
        fname_suffix=cmd_line_args['input_suffix'],
        extended_atom_types=cmd_line_args.get('extended_atom_types', False),
        model_task=model_task_)
    return checkpoint_path, model, model_kwargs, cmd_line_args, test_dl


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    parser.add_argument('model_checkpoint', type=str)
    parser.add_argument('test_types', type=str)
    parser.add_argument('test_data_root', type=str)
    parser.add_argument('--model_task', type=str,
                        help='(multitask models only) pose or affinity')
    parser.add_argument('--math', default='fp32',
                        choices=['fp32', 'bf16x3', 'bf16'],
                        help='arithmetic of the per-edge contractions')
    args = parser.parse_args(argv)
    checkpoint_path, model, _, _, test_dl = get_model_and_test_dl(
        Path(args.model_checkpoint).expanduser(), args.test_types,
        args.test_data_root, args.model_task)
    if args.model_task is not None:
        model.set_task({'pose': 'classification',
                        'affinity': 'regression'}[args.model_task])
    model.set_math(args.math)
    results_fname = Path(
        checkpoint_path.parents[1], 'predictions_{0}-{1}.txt'.format(
            Path(args.test_types).with_suffix('').name,
            checkpoint_path.with_suffix('').name)).expanduser()
    model.eval().val(test_dl, results_fname)
    return results_fname


if __name__ == '__main__':
    main()
