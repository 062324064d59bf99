"""Row N2 on the GPU: types file + parquets -> packed batch -> K1 -> scores,
against what the reference's own data loader returned for the same files
(tests/golden/loader.npz) and the CPU oracle run on those reference graphs."""
import numpy as np
import pytest
import torch
import yaml

from tests import helpers
from tests import gpu_helpers as gh
from tests.golden.loader_configs import CONFIGS, ROOT
from oracle import radius_graph as rg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold():
    return helpers.load_npz('loader.npz')


def _reference_batch(gold, cfg, items):
    """Concatenate the reference loader's per-complex graphs as PyG would."""
    xs, ps, rows, cols, attrs, batch, off = [], [], [], [], [], [], 0
    for b, i in enumerate(items):
        x = gold[f'{cfg}/{i}/x'].astype(np.float32)
        ei = gold[f'{cfg}/{i}/edge_index'].astype(np.int64)
        xs.append(x)
        ps.append(gold[f'{cfg}/{i}/pos'])
        rows.append(ei[0] + off)
        cols.append(ei[1] + off)
        attrs.append(gold[f'{cfg}/{i}/edge_attr'])
        batch.append(np.full(len(x), b))
        off += len(x)
    return (np.concatenate(xs), np.concatenate(ps), np.concatenate(rows),
            np.concatenate(cols), np.concatenate(attrs), np.concatenate(batch))


@pytest.mark.parametrize('cfg', sorted(CONFIGS))
@pytest.mark.parametrize('workers,device_crop', [(0, False), (3, False),
                                                 (0, True), (3, True)])
def test_packed_loader_graph_is_the_reference_graph(cfg, workers, device_crop,
                                                    gold):
    """Host crop (numpy) and device crop (K0) both reproduce the reference
    loader's nodes, features and edges bit for bit."""
    from pointvs_b200 import data
    ds = data.ComplexDataset(ROOT, types_fname=ROOT / 'pose.types', rot=False,
                             device='cuda', device_crop=device_crop,
                             **CONFIGS[cfg])
    n = len(ds)
    seen = 0
    for batch in data.PackedLoader(ds, batch_size=4, num_workers=workers):
        items = list(range(seen, min(n, seen + 4)))
        seen += len(items)
        x, pos, row, col, attr, _ = _reference_batch(gold, cfg, items)
        np.testing.assert_array_equal(batch.x.cpu().numpy(), x)
        np.testing.assert_array_equal(batch.pos.cpu().numpy(), pos)
        assert batch.num_graphs == len(items)
        order = rg.csr_order(row, col, attr)
        want = (row[order], col[order], attr[order])
        ei = batch.edge_index.cpu().numpy()
        np.testing.assert_array_equal(ei[0], want[0])
        np.testing.assert_array_equal(ei[1], want[1])
        np.testing.assert_array_equal(
            batch.edge_attr.argmax(dim=1).cpu().numpy(), want[2])
        np.testing.assert_array_equal(
            batch.y.numpy(), [int(gold[f'{cfg}/{i}/y']) for i in items])
        assert [str(p) for p in batch.lig_fname] == [
            str(gold[f'{cfg}/{i}/lig']) for i in items]
    assert seen == n


def _write_run(tmp_path, kw, cfg, model_type='egnn'):
    import pointvs_b200 as pv
    run = tmp_path / 'run'
    (run / 'checkpoints').mkdir(parents=True)
    torch.manual_seed(11)
    C = pv.MultitaskSatorrasEGNN if model_type == 'multitask' \
        else pv.SartorrasEGNN
    model = C(run, 1e-3, 1e-4, None, None, silent=True, **kw)
    with open(run / 'model_kwargs.yaml', 'w') as f:
        yaml.dump(kw, f)
    c = CONFIGS[cfg]
    with open(run / 'cmd_args.yaml', 'w') as f:
        yaml.dump({'model': model_type, 'learning_rate': 1e-3,
                   'weight_decay': 1e-4, 'save_path': str(run),
                   'compact': c['compact'], 'radius': c['radius'],
                   'use_atomic_numbers': c['use_atomic_numbers'],
                   'hydrogens': c['polar_hydrogens'], 'batch_size': 3,
                   'edge_radius': c['edge_radius'],
                   'estimate_bonds': c.get('estimate_bonds', False),
                   'input_suffix': 'parquet', 'egnn_attention': True,
                   'model_task': 'classification'}, f)
    torch.save({'model_state_dict': model.state_dict(),
                'optimiser_state_dict': {}, 'p_epoch': 1},
               run / 'checkpoints' / 'pose_ckpt_epoch_1.pt')
    return run, model


@pytest.mark.parametrize('cfg,math', [('smina_r10_e4', 'fp32'),
                                      ('atomic_h_r6_e3', 'bf16x3'),
                                      ('smina_wide_r7_bonds', 'fp32')])
def test_inference_cli_scores_match_oracle_on_reference_graphs(
        tmp_path, gold, cfg, math):
    """`python -m pointvs_b200.inference ckpt types root`: every line of the
    predictions file against the oracle evaluated on the graphs the REFERENCE
    loader produced for the same complexes."""
    from pointvs_b200 import inference
    dim = int(gold[f'{cfg}/feature_dim'])
    kw = dict(dim_input=dim, dim_output=1, k=48, num_layers=3,
              graphnorm=True, edge_attention=True, node_attention=True,
              residual=True, normalize=True, tanh=True)
    run, model = _write_run(tmp_path, kw, cfg)
    argv = [str(run), str(ROOT / 'pose.types'), str(ROOT), '--math', math]
    if cfg == 'smina_wide_r7_bonds':      # spawned loader workers
        argv += ['--workers', '2', '--worker_processes']
    out = inference.main(argv)
    pose = out.parent / ('pose_' + out.name)
    lines = pose.read_text().splitlines()
    n = int(gold[f'{cfg}/n'])
    assert len(lines) == n
    from oracle import egnn_oracle
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for start in range(0, n, 3):      # GraphNorm statistics are per batch
        items = list(range(start, min(n, start + 3)))
        x, pos, row, col, attr, batch = _reference_batch(gold, cfg, items)
        want, _ = egnn_oracle.model_forward(
            sd, torch.from_numpy(x), torch.from_numpy(np.stack([row, col])),
            torch.from_numpy(pos),
            torch.nn.functional.one_hot(
                torch.from_numpy(attr.astype(np.int64)), 3),
            torch.from_numpy(batch), num_layers=3, multitask=False,
            model_task='classification', **helpers.oracle_kwargs(kw))
        want = torch.sigmoid(want).reshape(-1).numpy()
        for j, i in enumerate(items):
            y_true, rest = lines[i].split(' | ')
            y_pred, rec, lig = rest.split(' ')
            assert float(y_true) == float(gold[f'{cfg}/{i}/y'])
            assert abs(float(y_pred) - want[j]) <= 6e-4 + 1e-4 * abs(want[j])
            assert lig == str(gold[f'{cfg}/{i}/lig'])
            assert rec == str(gold[f'{cfg}/{i}/rec'])


def test_regression_batches_and_rotation(gold):
    from pointvs_b200 import data
    dl = data.get_data_loader(
        ROOT, types_fname=ROOT / 'affinity.types', batch_size=4, mode='val',
        model_task='multi_regression', rot=True, seed=5, device='cuda',
        **CONFIGS['smina_r10_e4'])
    b0 = next(iter(dl))
    want = np.concatenate([gold[f'multi_regression/{i}/y'] for i in range(4)])
    np.testing.assert_array_equal(b0.y.numpy(), want)
    # positions are rotated per complex about its centroid image; the graph is
    # that of the unrotated structure
    x, pos, row, col, attr, batch = _reference_batch(
        gold, 'smina_r10_e4', range(4))
    assert not np.allclose(b0.pos.cpu().numpy(), pos)
    got = b0.pos.cpu().numpy()
    first = batch == 0
    d_ref = np.linalg.norm(pos[first][:, None] - pos[first][None], axis=-1)
    d_got = np.linalg.norm(got[first][:, None] - got[first][None], axis=-1)
    np.testing.assert_allclose(d_ref, d_got, atol=2e-4)
    assert b0.edge_index.shape[1] == len(row)


def test_device_crop_equals_host_crop_arrays():
    """K0 output (coords fp64, bp, features, offsets) against build_complex on
    the host, two receptors interleaved in one batch, small and large radii."""
    from pointvs_b200 import data
    for cfg, radius in (('smina_r10_e4', 10), ('atomic_h_r6_e3', 2.5),
                        ('atomic_h_r6_e3', 60.0), ('smina_wide_r7_bonds', 7)):
        kw = dict(CONFIGS[cfg], radius=radius)
        ds = data.ComplexDataset(ROOT, types_fname=ROOT / 'pose.types',
                                 rot=False, device='cuda', device_crop=True,
                                 **kw)
        items = [0, 1, 2, 4, 6, 3, 5]
        coords, bp, feats, cptr = ds.crop_on_device(items)
        host = [ds.load(i) for i in items]
        sizes = [len(c) for c in host]
        np.testing.assert_array_equal(np.diff(cptr), sizes)
        np.testing.assert_array_equal(
            coords.cpu().numpy(), np.concatenate([c.coords for c in host]))
        np.testing.assert_array_equal(
            bp.cpu().numpy(), np.concatenate([c.bp for c in host]))
        np.testing.assert_array_equal(
            feats.cpu().numpy(), np.concatenate([c.feats for c in host]))


def test_device_crop_long_ligand_and_empty_batch():
    """A 'ligand' longer than the shared-memory staging chunk (the receptor
    file used as the ligand), and an empty batch."""
    from pointvs_b200 import data
    ds = data.ComplexDataset(ROOT, types_fname=ROOT / 'pose.types', rot=False,
                             device='cuda', device_crop=True,
                             **CONFIGS['smina_r10_e4'])
    coords, bp, feats, cptr = ds.crop_on_device([])
    assert coords.shape[0] == 0 and list(cptr) == [0]
    rec = data.read_structure(ROOT / 'rec_0.parquet')
    big = {k: v[:700].copy() for k, v in rec.items()}
    big['bp'][:] = 0
    emit, code = data.atom_codes(big, ds.n_features, False, False, False, None)
    lig = data.Ligand(np.stack([big['x'], big['y'], big['z']], 1), emit, code)
    coords, bp, feats, cptr = ds.crop_on_device([6], [lig])
    want = data.build_complex(rec, big, ds.n_features, ds.radius)
    np.testing.assert_array_equal(coords.cpu().numpy(), want.coords)
    np.testing.assert_array_equal(bp.cpu().numpy(), want.bp)
    np.testing.assert_array_equal(feats.cpu().numpy(), want.feats)


def test_training_on_packed_loader_matches_oracle_autograd(tmp_path, gold):
    """train_model over the packed loader (device crop -> K1 -> K2 -> K3 ->
    clip -> Adam) against the same three steps done on the CPU with the oracle
    forward, torch autograd and torch Adam on the reference loader's graphs."""
    import pointvs_b200 as pv
    from oracle import egnn_oracle
    from pointvs_b200 import data
    cfg = 'smina_r10_e4'
    kw = dict(dim_input=int(gold[f'{cfg}/feature_dim']), dim_output=1, k=32,
              num_layers=3, graphnorm=False, edge_attention=True,
              node_attention=True, residual=True, normalize=True, tanh=True)
    torch.manual_seed(3)
    model = pv.SartorrasEGNN(tmp_path, 2e-3, 1e-4, None, None, silent=True,
                             **kw).cuda()
    with torch.no_grad():          # make the coordinate path count
        for n, p in model.named_parameters():
            if n.endswith('coord_mlp.2.weight'):
                p.mul_(300.0)
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point())
          for k, v in model.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=2e-3, weight_decay=1e-4)
    n = int(gold[f'{cfg}/n'])
    want = []
    for start in range(0, n, 3):
        items = list(range(start, min(n, start + 3)))
        x, pos, row, col, attr, batch = _reference_batch(gold, cfg, items)
        out, _ = egnn_oracle.model_forward(
            sd, torch.from_numpy(x), torch.from_numpy(np.stack([row, col])),
            torch.from_numpy(pos),
            torch.nn.functional.one_hot(
                torch.from_numpy(attr.astype(np.int64)), 3),
            torch.from_numpy(batch), num_layers=3, multitask=False,
            model_task='classification', **helpers.oracle_kwargs(kw))
        y = torch.tensor([float(gold[f'{cfg}/{i}/y']) for i in items])
        loss = torch.nn.functional.binary_cross_entropy_with_logits(
            out.reshape(-1), y)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_value_(params, 1.0)
        opt.step()
        want.append(float(loss))
    dl = data.get_data_loader(
        ROOT, types_fname=ROOT / 'pose.types', batch_size=3, mode='val',
        rot=False, device='cuda', device_crop=True, **CONFIGS[cfg])
    model.only_save_best_models = True
    got = model.train_model(dl, epochs=1)
    assert len(got) == len(want) == 3
    np.testing.assert_allclose(got, want, rtol=2e-4)
    for k, v in model.state_dict().items():
        if v.is_floating_point():
            ref = sd[k].detach()
            assert helpers.scaled_err(v.cpu().numpy(), ref.numpy()) < 5e-3, k


def test_train_cli_pose_then_affinity_then_inference(tmp_path):
    """`python -m pointvs_b200.train multitask ...` (mirror of point_vs.py):
    pose epochs -> pose validation -> affinity epochs -> affinity validation
    on the sample complexes, the reference's output files, and the run
    directory loads back through load_model / inference."""
    from pointvs_b200 import inference, train
    from pointvs_b200.load_model import load_model
    run = tmp_path / 'run'
    model = train.main([
        'multitask', str(run), '--model_task', 'both',
        '--train_data_root_pose', str(ROOT),
        '--train_types_pose', str(ROOT / 'pose.types'), '-ep', '2',
        '--train_data_root_affinity', str(ROOT),
        '--train_types_affinity', str(ROOT / 'affinity.types'), '-ea', '1',
        '--test_data_root_pose', str(ROOT),
        '--test_types_pose', str(ROOT / 'pose.types'),
        '--test_data_root_affinity', str(ROOT),
        '--test_types_affinity', str(ROOT / 'affinity.types'),
        '--layers', '2', '-k', '32', '-b', '4', '--egnn_attention',
        '--node_attention', '--egnn_residual', '--egnn_normalise',
        '--egnn_tanh', '--multi_target_affinity', '--math', 'bf16x3',
        '--end_flag'])
    for name in ('cmd_args.yaml', 'model_kwargs.yaml', '_FINISHED',
                 'checkpoints/pose_ckpt_epoch_2.pt',
                 'checkpoints/affinity_ckpt_epoch_1.pt',
                 'pose_predictions.txt', 'affinity_predictions.txt'):
        assert (run / name).is_file(), name
    pose = (run / 'pose_predictions.txt').read_text().splitlines()
    assert len(pose) == 7 and all(' | ' in ln for ln in pose)
    aff = (run / 'affinity_predictions.txt').read_text().splitlines()
    assert len(aff) >= 7 and aff[0].rsplit(' | ', 1)[1] in ('pki', 'pkd', 'ic50')
    kwargs = yaml.safe_load((run / 'model_kwargs.yaml').read_text())
    assert kwargs['dim_input'] == 22 and kwargs['dim_output'] == 3
    # the run directory is a reference-format checkpoint
    path, loaded, _, cmd = load_model(run, model_task='affinity')
    assert path.name == 'affinity_ckpt_epoch_1.pt' and cmd['model'] == 'multitask'
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), loaded.state_dict()[k].cpu()), k
    out = inference.main([str(run), str(ROOT / 'pose.types'), str(ROOT),
                          '--model_task', 'pose', '--math', 'bf16x3'])
    lines = (out.parent / ('pose_' + out.name)).read_text().splitlines()
    assert len(lines) == 7


def test_resume_training_continues_from_the_checkpoint_epoch(tmp_path):
    """resume_training.py: the recorded flags rebuild the loaders and
    train_model starts at the checkpoint's epoch."""
    from pointvs_b200 import resume_training, train
    run = tmp_path / 'run'
    train.main(['egnn', str(run), '--train_data_root_pose', str(ROOT),
                '--train_types_pose', str(ROOT / 'pose.types'), '-ep', '1',
                '--test_data_root_pose', str(ROOT),
                '--test_types_pose', str(ROOT / 'pose.types'),
                '--layers', '2', '-k', '16', '-b', '4', '--egnn_attention',
                '--compact', '--use_atomic_numbers', '--hydrogens',
                '--radius', '6', '--edge_radius', '3'])
    assert (run / 'checkpoints' / 'pose_ckpt_epoch_1.pt').is_file()
    model = resume_training.main([str(run), '-ep', '3'])
    assert model.p_epoch == 3
    for epoch in (2, 3):
        assert (run / 'checkpoints' / f'pose_ckpt_epoch_{epoch}.pt').is_file()
    assert len((run / 'pose_predictions.txt').read_text().splitlines()) == 7
