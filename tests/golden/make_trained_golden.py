"""Trained-checkpoint golden (BASELINE configs[0], README.md:54-66 of the reference).

Run in the build container only (needs /root/reference, ~2 min on 8 vCPU):

    python tests/golden/make_trained_golden.py

1. Runs the reference's OWN training CLI, unmodified, through oracle/ref_shim.py:
       point_vs.py multitask <run> --model_task both -ea 1 -ep 1 --layers 3
       (pose: data/small_chembl_test, affinity: data/multi_classification_sample)
   on the CPU.  It trains one pose epoch and one affinity epoch and writes
   checkpoints/{pose,affinity}_ckpt_epoch_1.pt, model_kwargs.yaml, cmd_args.yaml
   and {pose,affinity}_predictions.txt.
2. Copies into tests/golden/trained/: the two checkpoints (weights only -- the
   Adam moments triple the size and are not on the scoring path), the two yaml
   files, a SUBSET of the evaluation complexes (their parquets and types lines),
   the lines the reference wrote for them, and -- from the reference model
   re-loaded from each checkpoint -- the raw logits of the subset (logits.npz).

tests/test_gpu_trained.py then loads the checkpoints through
pointvs_b200.load_model / inference and must reproduce the prediction lines to
the printed three decimals and the logits to 1e-4 (fp32 and bf16x3).
Nothing from this repository's CUDA path or oracle takes part here.
"""
import os
import runpy
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import yaml

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

from oracle import ref_shim  # noqa: E402

OUT = HERE / 'trained'
N_POSE, N_AFF = 16, 8


def run_reference_cli(run_dir):
    ref_shim.import_reference()          # shim + CPU device
    ref = Path(ref_shim.REFERENCE_ROOT)
    argv = ['point_vs.py', 'multitask', str(run_dir), '--model_task', 'both',
            '-ea', '1', '-ep', '1', '--layers', '3',
            '--train_data_root_pose', 'data/small_chembl_test',
            '--train_types_pose', 'data/small_chembl_test.types',
            '--train_data_root_affinity', 'data/multi_classification_sample',
            '--train_types_affinity', 'data/multi_classification_sample.types',
            '--test_data_root_pose', 'data/small_chembl_test',
            '--test_types_pose', 'data/small_chembl_test.types',
            '--test_types_affinity', 'data/multi_classification_sample.types',
            '--test_data_root_affinity', 'data/multi_classification_sample']
    cwd, old_argv = os.getcwd(), sys.argv
    os.chdir(ref)
    sys.argv = argv
    real_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    torch.manual_seed(0)
    np.random.seed(0)
    try:
        import point_vs.global_objects as go
        import point_vs.preprocessing.data_loaders as dl
        go.NUM_WORKERS = 0               # the shim's DataLoader: in-process
        dl.NUM_WORKERS = 0
        runpy.run_path(str(ref / 'point_vs.py'), run_name='__main__')
    finally:
        os.chdir(cwd)
        sys.argv = old_argv
        torch.cuda.is_available = real_avail


def pick_subset(types_file, pred_file, n):
    """First n lines of the types file whose (rec, lig) appear in the
    reference's predictions file, with the prediction line of each."""
    preds = {}
    for line in Path(pred_file).read_text().splitlines():
        parts = line.split()
        preds[(parts[3], parts[4])] = line
    chosen = []
    for line in Path(types_file).read_text().splitlines():
        parts = line.split()
        rec, lig = parts[3], parts[4]
        key = next((k for k in preds
                    if k[0].endswith(Path(rec).name) and k[1].endswith(Path(lig).name)
                    and Path(k[1]).parent.name == Path(lig).parent.name), None)
        if key is not None:
            chosen.append((line, preds[key], rec, lig))
        if len(chosen) == n:
            break
    return chosen


def reference_logits(run_dir, task, types_file, data_root):
    """Raw outputs of the reference model re-loaded from the checkpoint, on
    the subset, through the reference's own loader (batch size 4)."""
    from point_vs.models.load_model import load_model
    from point_vs.preprocessing.data_loaders import (
        PygPointCloudDataset, get_data_loader)
    ckpt = Path(run_dir, 'checkpoints', f'{task}_ckpt_epoch_1.pt')
    _, model, _, cmd = load_model(ckpt, silent=True)
    model = model.eval()
    model.set_task({'pose': 'classification', 'affinity': 'regression'}[task])
    dl = get_data_loader(
        Path(data_root), dataset_class=PygPointCloudDataset, batch_size=4,
        compact=cmd.get('compact', False), radius=cmd.get('radius', 10),
        use_atomic_numbers=cmd.get('use_atomic_numbers', False), rot=False,
        augmented_actives=0, min_aug_angle=0,
        polar_hydrogens=cmd.get('hydrogens', False), receptors=None,
        mode='val', types_fname=Path(types_file), fname_suffix='parquet',
        edge_radius=cmd.get('edge_radius', 4),
        estimate_bonds=cmd.get('estimate_bonds', False),
        model_task={'pose': 'classification', 'affinity': 'regression'}[task])
    outs = []
    with torch.no_grad():
        for graph in torch.utils.data.DataLoader(
                dl.dataset, batch_size=4, shuffle=False,
                collate_fn=ref_shim._collate):
            outs.append(model(graph).reshape(-1).double().numpy())
    return np.concatenate(outs)


def main():
    ref = Path('/root/reference')
    run_dir = Path(tempfile.mkdtemp(prefix='pvs_trained_')) / 'run'
    run_reference_cli(run_dir)
    if OUT.exists():
        shutil.rmtree(OUT)
    (OUT / 'run' / 'checkpoints').mkdir(parents=True)
    for name in ('model_kwargs.yaml', 'cmd_args.yaml'):
        shutil.copy(run_dir / name, OUT / 'run' / name)
    for task in ('pose', 'affinity'):
        ck = torch.load(run_dir / 'checkpoints' / f'{task}_ckpt_epoch_1.pt',
                        map_location='cpu', weights_only=False)
        ck.pop('optimiser_state_dict', None)
        torch.save(ck, OUT / 'run' / 'checkpoints' / f'{task}_ckpt_epoch_1.pt')
    logits = {}
    for task, n, types, root in (
            ('pose', N_POSE, 'data/small_chembl_test.types', 'data/small_chembl_test'),
            ('affinity', N_AFF, 'data/multi_classification_sample.types',
             'data/multi_classification_sample')):
        subset = pick_subset(ref / types, run_dir / f'{task}_predictions.txt', n)
        assert len(subset) == n, (task, len(subset))
        droot = OUT / f'data_{task}'
        for _, _, rec, lig in subset:
            for rel in (rec, lig):
                dst = droot / rel
                dst.parent.mkdir(parents=True, exist_ok=True)
                if not dst.exists():
                    shutil.copy(ref / root / rel, dst)
        (OUT / f'{task}.types').write_text(
            '\n'.join(s[0] for s in subset) + '\n')
        (OUT / f'{task}_predictions.txt').write_text(
            '\n'.join(s[1] for s in subset) + '\n')
        cwd = os.getcwd()
        os.chdir(ref)
        try:
            logits[task] = reference_logits(run_dir, task, OUT / f'{task}.types',
                                            droot)
        finally:
            os.chdir(cwd)
    np.savez_compressed(OUT / 'logits.npz', **logits)
    kw = yaml.safe_load((OUT / 'run' / 'model_kwargs.yaml').read_text())
    size = sum(f.stat().st_size for f in OUT.rglob('*') if f.is_file())
    print('trained golden:', {k: v.shape for k, v in logits.items()},
          'model_kwargs', kw, f'{size / 1e6:.2f} MB')


if __name__ == '__main__':
    main()
