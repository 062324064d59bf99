"""Generate the data-loader goldens (row N2) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_loader_golden.py

1. copies a handful of the reference's own sample complexes (parquet DATA
   files, a few hundred kB) to tests/golden/complexes/ and writes the types
   files that index them;
2. runs the reference's `PygPointCloudDataset` (imported through
   oracle/ref_shim.py) over them in several configurations and stores what it
   returns per complex -- node features, positions, edge list, edge classes,
   label -- in tests/golden/loader.npz.

pointvs_b200/data.py is compared against these arrays on the CPU
(tests/test_data_cpu.py) and, through K1, on the GPU (tests/test_gpu_data.py).
"""
import shutil
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

from oracle import ref_shim  # noqa: E402

REF = ref_shim.import_reference()
REFROOT = Path(ref_shim.REFERENCE_ROOT)
OUT = HERE / 'complexes'

CLASSIFICATION = [
    (1, 'receptors/18061.parquet', 'ligands/18061_actives/mol48_0.parquet'),
    (1, 'receptors/10378.parquet', 'ligands/10378_actives/mol7_5.parquet'),
]

from tests.golden.loader_configs import CONFIGS  # noqa: E402


def copy_inputs():
    src = REFROOT / 'data' / 'small_chembl_test'
    lines = (REFROOT / 'data' / 'small_chembl_test.types').read_text(
        ).splitlines()
    chosen = list(CLASSIFICATION)
    for rec in ('receptors/18061.parquet', 'receptors/10378.parquet'):
        decoys = [ln.split() for ln in lines
                  if ln.split()[3] == rec and ln.split()[0] == '0']
        chosen += [(0, d[3], d[4]) for d in decoys[:2]]
    if OUT.exists():
        shutil.rmtree(OUT)
    for _, rec, lig in chosen:
        for rel in (rec, lig):
            dst = OUT / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            if not dst.exists():
                shutil.copyfile(src / rel, dst)
                dst.chmod(0o644)
    # the reference's own unit-test complex
    for name in ('rec_0.parquet', 'lig_0.parquet'):
        shutil.copyfile(REFROOT / 'test' / 'resources' / name, OUT / name)
        (OUT / name).chmod(0o644)
    chosen.append((1, 'rec_0.parquet', 'lig_0.parquet'))
    with open(OUT / 'pose.types', 'w') as f:
        for i, (label, rec, lig) in enumerate(chosen):
            if i == 3:
                f.write('\n')      # blank lines are skipped by the reference
            f.write(f'{label} -1 {-1.0 if i % 2 else 0.5 * i} {rec} {lig}\n')
    rng = np.random.default_rng(0)
    with open(OUT / 'affinity.types', 'w') as f:
        for label, rec, lig in chosen:
            vals = np.round(rng.uniform(-1, 9, size=3), 3)
            f.write(f'{vals[0]} {vals[1]} {vals[2]} {rec} {lig}\n')
        # a complex whose ligand file does not exist is dropped
        f.write('1.0 2.0 3.0 rec_0.parquet ligands/missing.parquet\n')
    with open(OUT / 'nolabel.types', 'w') as f:
        for label, rec, lig in chosen[:3]:
            f.write(f'{rec} {lig}\n')
    return chosen


def run_reference(types_name, model_task, **kw):
    from point_vs.preprocessing.data_loaders import PygPointCloudDataset
    ds = PygPointCloudDataset(
        OUT, types_fname=OUT / types_name, model_task=model_task, rot=False,
        fname_suffix='parquet', **kw)
    items = []
    for i in range(len(ds)):
        d = ds[i]
        items.append(dict(
            x=d.x.numpy().astype(np.uint8),
            pos=d.pos.numpy().astype(np.float32),
            edge_index=d.edge_index.numpy().astype(np.int32),
            edge_attr=d.edge_attr.argmax(dim=1).numpy().astype(np.uint8),
            y=d.y.numpy(),
            lig=str(d.lig_fname), rec=str(d.rec_fname)))
        assert torch.equal(d.x, torch.from_numpy(items[-1]['x']).float())
    return ds, items


def main():
    copy_inputs()
    out = {}
    for name, kw in CONFIGS.items():
        ds, items = run_reference('pose.types', 'classification', **kw)
        out[f'{name}/n'] = np.int64(len(items))
        out[f'{name}/feature_dim'] = np.int64(ds.feature_dim)
        for i, it in enumerate(items):
            for k, v in it.items():
                out[f'{name}/{i}/{k}'] = np.asarray(v)
        print(name, 'items', len(items), 'feature_dim', ds.feature_dim,
              'atoms', [len(it['x']) for it in items],
              'edges', [it['edge_index'].shape[1] for it in items])
    for task in ('regression', 'multi_regression'):
        ds, items = run_reference('affinity.types', task,
                                  **CONFIGS['smina_r10_e4'])
        out[f'{task}/n'] = np.int64(len(items))
        for i, it in enumerate(items):
            out[f'{task}/{i}/y'] = it['y']
            out[f'{task}/{i}/lig'] = np.asarray(it['lig'])
        print(task, [it['y'].tolist() for it in items])
    # sampler weights of the classification set
    ds, _ = run_reference('pose.types', 'classification',
                          **CONFIGS['smina_r10_e4'])
    out['sample_weights'] = ds.sample_weights.numpy()
    out['labels'] = np.asarray(ds.labels, dtype=np.int64)
    np.savez_compressed(HERE / 'loader.npz', **out)
    print('loader.npz written:', (HERE / 'loader.npz').stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
