"""pipeline.ScoreStream: pipelined host-fed scoring returns exactly what
scoring the same batches one at a time returns, in order."""
import numpy as np
import pytest
import torch

from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

KW = dict(dim_input=13, dim_output=1, k=64, num_layers=3, graphnorm=False,
          edge_attention=True, node_attention=True, residual=True,
          normalize=True, tanh=True)


def _batches():
    from pointvs_b200.synthetic import synthetic_batch
    out = []
    # sizes change between submits so the staging slots have to grow
    for seed, b, atoms in ((1, 3, 200), (2, 6, 350), (3, 2, 120), (4, 6, 360),
                           (5, 1, 90), (6, 4, 300), (7, 5, 340)):
        out.append(synthetic_batch(100 * seed, b, atoms, 12, ragged=True))
    return out


@pytest.mark.parametrize('math,capacity', [('fp32', 'auto'), ('bf16x3', None),
                                           ('fp16x2', 'auto')])
def test_stream_equals_one_at_a_time(math, capacity):
    import pointvs_b200 as pv
    from pointvs_b200.pipeline import ScoreStream
    model = gh.build_model(KW, seed=2, coord_gain=1.0)
    model.set_math(math)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    batches = _batches()
    want = []
    with torch.no_grad():
        for coords, bp, feats, cptr in batches:
            b = pv.PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0)
            want.append(model(b).reshape(len(cptr) - 1, -1).cpu().numpy())
    stream = ScoreStream(model, 4.0, 4.0, depth=2, edge_capacity=capacity)
    got = []
    for i, (coords, bp, feats, cptr) in enumerate(batches):
        if i % 2:      # pinned tensors on odd submits, plain numpy on even
            coords = torch.from_numpy(coords).pin_memory()
            feats = torch.from_numpy(feats).pin_memory()
        stream.submit(coords, bp, feats, cptr, tag=i)
        got += stream.results()
    got += stream.drain()
    assert [t for t, _ in got] == list(range(len(batches)))
    for (_, g), w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_stream_reports_edge_overflow():
    from pointvs_b200.pipeline import ScoreStream
    model = gh.build_model(KW, seed=2)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    coords, bp, feats, cptr = _batches()[1]
    stream = ScoreStream(model, 4.0, 4.0, edge_capacity=100)
    stream.submit(coords, bp, feats, cptr)
    with pytest.raises(RuntimeError, match='edge capacity'):
        stream.drain()


def test_screen_uses_the_stream_and_matches_direct_scoring():
    import pointvs_b200 as pv
    from pointvs_b200 import parallel
    from pointvs_b200.synthetic import synthetic_complex
    model = gh.build_model(KW, seed=4)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    complexes = [synthetic_complex(s, 150 + 10 * s, 10) for s in range(9)]
    got = parallel.screen(model, complexes, batch_size=4, activation=None)
    with torch.no_grad():
        for i, (coords, bp, feats) in enumerate(complexes):
            b = pv.PackedBatch.from_arrays(coords, bp, feats, [0, len(bp)],
                                           4.0, 4.0)
            want = model(b).reshape(-1).cpu().numpy()
            np.testing.assert_allclose(got[i].reshape(-1), want, rtol=2e-5,
                                       atol=1e-6)
