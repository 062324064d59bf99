"""CPU-side checks of the C ABI: the library loads, exports every function
include/pvs_b200.h declares, validates arguments before touching the GPU, and
the host-side mirror keeps the reference's state_dict layout."""
import ctypes as C
import os
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared_functions():
    text = (ROOT / 'include' / 'pvs_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pvs_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from pointvs_b200 import _cabi
    lib = _cabi.lib()
    names = _declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'declared in pvs_b200.h but not exported: {missing}'


def test_version_capabilities_and_status_strings():
    from pointvs_b200 import _cabi
    lib = _cabi.lib()
    assert lib.pvs_version() == 100
    caps = lib.pvs_capabilities()
    assert caps & _cabi.CAP_FWD_FP32 and caps & _cabi.CAP_FWD_TCGEN05
    assert caps & _cabi.CAP_BWD_FP32
    assert lib.pvs_status_string(0) == b'ok'
    assert b'hidden width' in lib.pvs_status_string(2)


def test_argument_validation_happens_before_any_launch():
    from pointvs_b200 import _cabi
    lib = _cabi.lib()
    # k out of range -> workspace query reports -1, layer call reports status 2
    cfg = _cabi.LayerConfig(96, 3, 0, 0, 0, 0, None, None)
    assert lib.pvs_egnn_layer_workspace_bytes(10, 10, C.byref(cfg)) == -1
    g = _cabi.Graph(10, 10, None, None, None, None, None, 1)
    p = _cabi.LayerParams()
    rc = lib.pvs_egnn_layer_fwd(C.byref(g), C.byref(cfg), C.byref(p), None,
                                None, None, None, None, None, None, None, None,
                                C.c_int64(0), None)
    assert rc == 2
    cfg = _cabi.LayerConfig(64, 3, 0, 0, 0, 0, None, None)
    rc = lib.pvs_egnn_layer_fwd(C.byref(g), C.byref(cfg), C.byref(p), None,
                                None, None, None, None, None, None, None, None,
                                C.c_int64(0), None)
    assert rc == 1          # null pointers
    assert lib.pvs_linear_fwd(None, 4, 3, 4, None, 4, None, 300, 0, None, 300,
                              None) == 1
    assert lib.pvs_radius_graph_count(
        None, None, None, -1, 0, 0, C.c_double(4.0), C.c_double(2.0), None,
        None, None, None, None) == 1


def test_structs_match_header_layout():
    from pointvs_b200 import _cabi
    assert C.sizeof(_cabi.LayerConfig) == 6 * 4 + 3 * 8
    assert C.sizeof(_cabi.LayerParams) == 20 * 8
    assert C.sizeof(_cabi.LayerGrads) == 20 * 8
    assert C.sizeof(_cabi.Graph) == 8 + 5 * 8 + 8 + 2 * 8 + 8


def test_missing_library_fails_loudly(monkeypatch):
    from pointvs_b200 import _cabi
    monkeypatch.setattr(_cabi, '_lib', None)
    monkeypatch.setattr(_cabi, 'LIB_PATH', '/nonexistent/libpvs_b200.so')
    with pytest.raises(_cabi.PvsError, match='no CPU fallback'):
        _cabi.lib()


def test_state_dict_layout_and_init_match_reference_goldens():
    """Keys/shapes equal the reference's (checkpoint compatibility), and with
    the same seed the modules draw the same initial weights as the reference
    did when the golden was generated (same creation order)."""
    import pointvs_b200 as pv
    from tests import helpers
    for name, (cls, kw, _) in helpers.MODEL_GOLDENS.items():
        _, sd = helpers.load_model_golden(name)
        klass = pv.SartorrasEGNN if cls == 'egnn' else pv.MultitaskSatorrasEGNN
        model = klass(Path('/tmp/pvs_test'), 0, 0, None, None, silent=True, **kw)
        own = model.state_dict()
        assert set(own) == set(sd), name
        for k in sd:
            assert tuple(own[k].shape) == tuple(sd[k].shape), (name, k)
    torch.manual_seed(0)
    cls, kw, _ = helpers.MODEL_GOLDENS['cfg3_k64_l8']
    model = pv.SartorrasEGNN(Path('/tmp/pvs_test'), 0, 0, None, None,
                             silent=True, **kw)
    _, sd = helpers.load_model_golden('cfg3_k64_l8')
    assert model.param_count == 236497
    # golden was made with coord gain x1000; every other tensor is bit-equal
    for k, v in model.state_dict().items():
        if k.endswith('coord_mlp.2.weight'):
            assert torch.allclose(v.cpu() * 1000.0, sd[k], rtol=1e-6)
        else:
            assert torch.equal(v.cpu(), sd[k]), k


def test_cli_flag_semantics_multitask_attention_placement():
    import pointvs_b200 as pv
    kw = dict(dim_input=13, dim_output=1, k=16, num_layers=4,
              edge_attention=True, node_attention=True,
              edge_attention_first_only=True, node_attention_final_only=True,
              graphnorm=False)
    m = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_test'), 0, 0, None, None,
                                 silent=True, **kw)
    layers = list(m.layers)[1:]
    assert [l.edge_attention for l in layers] == [True, False, False, False]
    assert [l.node_attention for l in layers] == [False, False, False, True]
    with pytest.raises(ValueError):
        m.set_task('ranking')


def test_entry_points_fail_loudly_without_a_gpu_path(tmp_path):
    """Training / streaming entry points refuse what is outside the path
    before any kernel is involved, and never fall back to the CPU."""
    import pytest
    import torch
    from pointvs_b200 import train
    from pointvs_b200._cabi import PvsError
    with pytest.raises(NotImplementedError):
        train.main(['lucid', str(tmp_path / 'a')])
    with pytest.raises(RuntimeError):
        train.main(['egnn', str(tmp_path / 'b'), '--model_task', 'both'])
    with pytest.raises(NotImplementedError):
        train.main(['egnn', str(tmp_path / 'c'), '--double'])
    if not torch.cuda.is_available():
        import pointvs_b200 as pv
        from pointvs_b200.pipeline import ScoreStream
        model = pv.SartorrasEGNN(tmp_path, 0, 0, None, None, silent=True,
                                 dim_input=13, dim_output=1, k=16,
                                 num_layers=1)
        with pytest.raises(PvsError):
            ScoreStream(model)
