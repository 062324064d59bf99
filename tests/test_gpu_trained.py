"""Trained-checkpoint parity (BASELINE configs[0]; SURVEY.md 8d "Parity use").

tests/golden/trained/ holds what the reference's own training CLI produced
(README.md:54-66 of the reference, run unmodified through oracle/ref_shim.py
by tests/golden/make_trained_golden.py): the pose and affinity checkpoints of
a 3-layer multitask EGNN after one epoch each, the yaml files next to them, a
subset of the evaluation complexes, the prediction lines the reference wrote
for that subset and the raw logits of the reference model re-loaded from each
checkpoint.

The CUDA path loads the checkpoints through load_model / inference and must
reproduce the lines to the printed three decimals and the logits within
1e-4 relative (+ 2e-6 absolute: a logit is a difference of O(1) terms), in the
fp32 FFMA mode and in the default tcgen05 bf16x3 mode.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).resolve().parent / 'golden' / 'trained'
TASKS = {'pose': 'classification', 'affinity': 'regression'}


def _gold_lines(task):
    return (GOLD / f'{task}_predictions.txt').read_text().splitlines()


def _model_and_loader(run, task):
    """What the reference's training CLI does for its final validation pass
    (point_vs.py:263-271): the run's own loader settings with the task's
    model_task.  (Its inference.py would turn `--model_task affinity` on this
    run into multi_regression -- inference.py:49-55 -- which a dim_output = 1
    model cannot serve; pointvs_b200.inference mirrors that, so the affinity
    half of this test builds the loader itself.)"""
    from pointvs_b200 import data
    from pointvs_b200.load_model import load_model
    _, model, _, cmd = load_model(run, model_task=task)
    model.set_task(TASKS[task])
    dl = data.get_data_loader(
        str(GOLD / f'data_{task}'), receptors=None,
        compact=cmd.get('compact', False),
        use_atomic_numbers=cmd.get('use_atomic_numbers', False),
        radius=cmd.get('radius', 10),
        polar_hydrogens=cmd.get('hydrogens', False), batch_size=4,
        types_fname=str(GOLD / f'{task}.types'),
        edge_radius=cmd.get('edge_radius', 4),
        estimate_bonds=cmd.get('estimate_bonds', False),
        prune=cmd.get('prune', False), rot=False, mode='val',
        fname_suffix=cmd.get('input_suffix', 'parquet'),
        extended_atom_types=cmd.get('extended_atom_types', False),
        model_task=TASKS[task], device_crop=True)
    return model, dl


def test_fixture_is_self_consistent():
    """(CPU) the stored logits reproduce the reference's printed predictions."""
    logits = np.load(GOLD / 'logits.npz')
    for task in TASKS:
        lines = _gold_lines(task)
        assert len(lines) == len(logits[task])
        for line, z in zip(lines, logits[task]):
            printed = float(line.split()[2])
            val = 1.0 / (1.0 + np.exp(-z)) if task == 'pose' else z
            assert abs(printed - val) <= 5.01e-4, (task, line, z)
    for task in TASKS:
        assert (GOLD / 'run' / 'checkpoints' / f'{task}_ckpt_epoch_1.pt').is_file()


@pytest.mark.gpu
@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('task', ['pose', 'affinity'])
def test_prediction_lines_match_the_reference_run(tmp_path, task, math):
    import shutil
    from pointvs_b200 import inference
    run = tmp_path / 'run'
    shutil.copytree(GOLD / 'run', run)
    if task == 'pose':          # through the scoring entry point
        out = inference.main([str(run), str(GOLD / f'{task}.types'),
                              str(GOLD / f'data_{task}'), '--model_task', task,
                              '--math', math, '--batch_size', '4'])
    else:
        model, dl = _model_and_loader(run, task)
        model.set_math(math)
        out = run / 'predictions.txt'
        model.eval().val(dl, out)
    pred = out.parent / (f'{task}_' + out.name)
    got = pred.read_text().splitlines()
    want = _gold_lines(task)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        gp, wp = g.split(), w.split()
        assert gp[3:] == wp[3:], (g, w)                  # receptor, ligand
        assert float(gp[0]) == float(wp[0])              # label
        # the printed value is rounded to 3 decimals: equal, or one unit in
        # the last place apart when the exact value sits on a rounding edge
        assert abs(float(gp[2]) - float(wp[2])) <= 1.001e-3, (g, w)
    same = sum(g.split()[2] == w.split()[2] for g, w in zip(got, want))
    assert same >= len(want) - 1, f'{same}/{len(want)} identical lines'


@pytest.mark.gpu
@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('task', ['pose', 'affinity'])
def test_logits_match_the_reference_model(task, math):
    want = np.load(GOLD / 'logits.npz')[task]
    model, dl = _model_and_loader(GOLD / 'run', task)
    model.set_math(math)
    got = []
    with torch.no_grad():
        for graph in dl:
            y_pred, _, _, _ = model.unpack_input_data_and_predict(graph)
            got.append(y_pred.reshape(-1).double().cpu().numpy())
    got = np.concatenate(got)
    assert got.shape == want.shape
    err = np.abs(got - want)
    bound = 1e-4 * np.abs(want) + 2e-6
    assert np.all(err <= bound), (float(err.max()), float((err / bound).max()))
