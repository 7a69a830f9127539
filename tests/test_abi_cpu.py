"""CPU: the C-ABI library builds, loads, and exports every symbol include/gp_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'gp_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gp_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from graph_pooling_b200 import _lib, build
    build.build_native()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'missing export: ' + s
    # and the python binding covers the same set
    assert sorted(_lib.EXPORTS) == syms


def test_version_and_error_accessor():
    from graph_pooling_b200 import _lib
    lib = _lib.load()
    assert lib.gp_version() >= 100
    assert isinstance(lib.gp_last_error(), bytes)
    # argument validation happens before any CUDA call: usable without a GPU
    assert lib.gp_bgemm_f32(None, None) == -1
    assert b'null' in lib.gp_last_error()


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly, not routed to an eager fallback."""
    import numpy as np
    import torch
    from graph_pooling_b200 import encoders
    m = encoders.GcnEncoderGraph(3, 8, 8, 2, 3)
    x = torch.zeros(2, 5, 3)
    adj = torch.zeros(2, 5, 5)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(x, adj, np.array([5, 5]))


def test_state_dict_keys_match_oracle():
    from graph_pooling_b200 import encoders
    from oracle import diffpool_oracle as orc
    for P in (1, 2):
        a = encoders.SoftPoolingGcnEncoder(40, 3, 30, 30, 6, 3, 30, assign_ratio=0.25, num_pooling=P)
        b = orc.SoftPoolingGcnEncoder(40, 3, 30, 30, 6, 3, 30, assign_ratio=0.25, num_pooling=P)
        assert [(k, tuple(v.shape)) for k, v in a.state_dict().items()] == \
               [(k, tuple(v.shape)) for k, v in b.state_dict().items()]
    a = encoders.GcnEncoderGraph(7, 20, 20, 2, 3)
    b = orc.GcnEncoderGraph(7, 20, 20, 2, 3)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    a = encoders.GcnSet2SetEncoder(7, 20, 20, 2, 3)
    b = orc.GcnSet2SetEncoder(7, 20, 20, 2, 3)
    assert [(k, tuple(v.shape)) for k, v in a.state_dict().items()] == \
           [(k, tuple(v.shape)) for k, v in b.state_dict().items()]


def test_host_side_plan_logic(monkeypatch):
    """Host logic that needs no GPU: dead-cluster padding of the cluster count, per-layer dropout seeds, and the
    dropout schedule of a stack (conv_block layers only, training mode only: encoders.py:1015, nn.Dropout)."""
    from graph_pooling_b200 import encoders, engine
    assert [encoders._kpad(k) for k in (8, 10, 62, 250, 512, 1250)] == [8, 16, 64, 256, 512, 1256]
    monkeypatch.setenv('GP_NO_KPAD', '1')
    assert encoders._kpad(250) == 250
    seeds = {engine.layer_seed(12345, l) for l in range(8)}
    assert len(seeds) == 8 and all(0 <= s < (1 << 62) for s in seeds)
    m = encoders.GcnEncoderGraph(7, 20, 20, 2, 4, dropout=0.25)
    assert m._drops(m.conv_block) == [0.0, 0.25, 0.25, 0.0]
    m.eval()
    assert m._drops(m.conv_block) == [0.0, 0.0, 0.0, 0.0]
    s = encoders.SoftPoolingGcnEncoder(40, 3, 16, 16, 2, 4, 16, dropout=0.5)
    assert s._drops(s.conv_block) == [0.0] * 4                      # R8: the first GCN ignores dropout (:1172-1173)
    assert s._drops(s.conv_block_after_pool[0]) == [0.0, 0.5, 0.5, 0.0]
    with pytest.raises(ValueError):
        encoders.Set2Set(10, 15)                                    # hidden_dim must be 2 * input_dim
