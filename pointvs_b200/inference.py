"""Scoring entry point, mirror of /root/reference/point_vs/inference.py.

    python -m pointvs_b200.inference <model_checkpoint> <test_types> \
        <test_data_root> [--model_task pose|affinity]

Loads a reference-format checkpoint into the CUDA-backed model, scores the
types file and writes `predictions_<types>-<ckpt>.txt` next to the checkpoint
in the reference's line format.  Complexes come from `pointvs_b200.data`
(types file + parquets -> packed batches whose radius graph is built on the
device); `--reference_loader` feeds the reference's own PyG data loader through
the same model instead, if PointVS is importable.
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
            '--reference_loader needs PointVS itself '
            '(point_vs.preprocessing.data_loaders) to be importable') from exc

    def factory(data_root, **kwargs):
        return get_data_loader(data_root, dataset_class=PygPointCloudDataset,
                               **kwargs)
    return factory


def _packed_loader_factory():
    from .data import get_data_loader   # noqa: PLC0415
    return get_data_loader


def get_model_and_test_dl(checkpoint_path, test_types, test_data_root,
                          model_task=None, loader_factory=None,
                          batch_size=None, device_crop=False,
                          loader_kwargs=None):
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
    factory = loader_factory or _packed_loader_factory()
    test_dl = factory(
        test_data_root, receptors=None,
        compact=cmd_line_args.get('compact', False),
        use_atomic_numbers=cmd_line_args.get('use_atomic_numbers', False),
        radius=cmd_line_args.get('radius', 10),
        polar_hydrogens=cmd_line_args.get('hydrogens', False),
        batch_size=batch_size or cmd_line_args.get('batch_size', 32),
        types_fname=test_types,
        edge_radius=cmd_line_args.get('edge_radius', 4),
        estimate_bonds=cmd_line_args.get('estimate_bonds', False),
        prune=cmd_line_args.get('prune', False), rot=False, mode='val',
        fname_suffix=cmd_line_args.get('input_suffix', 'parquet'),
        extended_atom_types=cmd_line_args.get('extended_atom_types', False),
        model_task=model_task_, **({'device_crop': True} if device_crop
                                   else {}), **(loader_kwargs or {}))
    return checkpoint_path, model, model_kwargs, cmd_line_args, test_dl


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    parser.add_argument('model_checkpoint', type=str)
    parser.add_argument('test_types', type=str)
    parser.add_argument('test_data_root', type=str)
    parser.add_argument('--model_task', type=str,
                        help='(multitask models only) pose or affinity')
    parser.add_argument('--math', default='bf16x3',
                        choices=['fp32', 'bf16x3', 'bf16', 'fp16x2'],
                        help='arithmetic of the per-edge contractions')
    parser.add_argument('--reference_loader', action='store_true',
                        help="use PointVS's own PyG data loader")
    parser.add_argument('--host_crop', action='store_true',
                        help='crop / type the complexes on the host instead '
                             'of on the device (K0)')
    parser.add_argument('--workers', type=int, default=4,
                        help='loader workers reading the parquets')
    parser.add_argument('--worker_processes', action='store_true',
                        help='loader workers are spawned processes instead of '
                             'threads (no interpreter lock: ~10x the rate)')
    parser.add_argument('--batch_size', type=int, default=None,
                        help='complexes per packed batch (default: the '
                             'value the model was trained with)')
    args = parser.parse_args(argv)
    checkpoint_path, model, _, _, test_dl = get_model_and_test_dl(
        Path(args.model_checkpoint).expanduser(), args.test_types,
        args.test_data_root, args.model_task,
        loader_factory=_reference_loader_factory()
        if args.reference_loader else None, batch_size=args.batch_size,
        device_crop=not (args.host_crop or args.reference_loader),
        loader_kwargs=None if args.reference_loader else dict(
            num_workers=args.workers, worker_processes=args.worker_processes))
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
