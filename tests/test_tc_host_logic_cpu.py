"""CPU: host-side decisions of the tensor-core schedule that need no GPU (engine_tc.py).

* pick_split: the split-K factor fills the persistent grid it runs on (74 CTA pairs or 148 CTAs) in at most three
  rounds, keeps at least four 64-wide k-steps per split, and never does worse than the round-1 rule's 2-wave target;
* bfbuf: row padding of operand buffers (8 elements for TMA, 32 for adjacency-sized operands of the row epilogue);
* upper_band_ok / chain_ok: the conditions under which the link loss keeps dl/dP as its upper band and the pooling
  contraction goes through the chained kernel."""
import os

import pytest
import torch

from graph_pooling_b200 import engine_tc as T


def _grid(M, N):
    t128, t256 = (M + 127) // 128, (M + 255) // 256
    pair = N > 128 and t128 >= 2 and 2 * t256 * 16 <= t128 * 17
    tiles = (t256 if pair else t128) * ((N + 255) // 256 if N > 128 else 1)
    return tiles, (74 if pair else 148)


@pytest.mark.parametrize('M,N,K', [(512, 768, 524288), (1250, 1506, 320000), (128, 128, 524288), (128, 512, 524288),
                                   (384, 128, 524288), (64, 16, 4096), (30, 30, 2000), (512, 384, 20000), (2048, 2048, 64)])
def test_pick_split_fills_the_grid(M, N, K, monkeypatch):
    monkeypatch.delenv('GP_NO_PAIR', raising=False)
    s = T.pick_split(M, N, K)
    tiles, units = _grid(M, N)
    assert s >= 1 and s <= max(1, K // 256)
    items = tiles * s
    rounds = -(-items // units)
    assert rounds <= 3 or s == 1
    # no other admissible factor fills the grid better
    eff = items / float(units * rounds)
    for t in range(1, min(max(1, K // 256), (3 * units) // tiles + 1) + 1):
        it = tiles * t
        r = -(-it // units)
        if r <= 3:
            assert it / float(units * r) <= eff + 1e-9


def test_pick_split_cfg4_dwp_is_exactly_three_rounds(monkeypatch):
    monkeypatch.delenv('GP_NO_PAIR', raising=False)
    assert T.pick_split(512, 768, 256 * 2048) == 37          # 6 pair tiles x 37 = 222 = 3 x 74
    monkeypatch.setenv('GP_NO_PAIR', '1')
    s = T.pick_split(512, 768, 256 * 2048)                   # single CTAs: 12 tiles on 148 CTAs
    assert (12 * s) % 148 in (0, 144) or 12 * s <= 148 * 3


def test_bfbuf_padding():
    ws = type('W', (), {'device': torch.device('cpu')})()
    a = T.bfbuf(ws, 2, 5, 30)
    assert (a.ld, a.sb, tuple(a.t.shape)) == (32, 5 * 32, (2, 5, 32))
    b = T.bfbuf(ws, 2, 5000, 5000, pad=32)
    assert b.ld == 5024 and b.ld % 16 == 0 and b.sb == 5000 * 5024
    c = T.bfbuf(ws, 1, 3, 2048, pad=32)
    assert c.ld == 2048


def test_upper_band_and_chain_conditions(monkeypatch):
    ws = type('W', (), {'device': torch.device('cpu')})()
    monkeypatch.delenv('GP_NO_UPPER_G', raising=False)
    monkeypatch.delenv('GP_CHAIN', raising=False)
    N, K = 300, 80
    sb, adj8, adj32 = T.bfbuf(ws, 2, N, K), T.bfbuf(ws, 2, N, N), T.bfbuf(ws, 2, N, N, pad=32)
    flags = object()
    assert T.upper_band_ok(sb, adj32, N, 0, flags)
    assert not T.upper_band_ok(sb, adj32, N, 1, flags)            # Frobenius option: full G
    assert not T.upper_band_ok(sb, adj32, N, 0, None)             # no adjacency flags: nothing known about symmetry
    assert not T.upper_band_ok(sb, adj8, N, 0, flags)             # rows not padded to 32 elements
    monkeypatch.setenv('GP_NO_UPPER_G', '1')
    assert not T.upper_band_ok(sb, adj32, N, 0, flags)
    # chained pooling is opt-in and limited to 512 clusters
    monkeypatch.setattr(T, 'CHAIN_POOLING', None)
    assert not T.chain_ok(sb, adj32, N, K)
    monkeypatch.setenv('GP_CHAIN', '1')
    assert T.chain_ok(sb, adj32, N, K)
    assert not T.chain_ok(sb, adj32, N, 520)
    monkeypatch.setattr(T, 'CHAIN_POOLING', False)
    assert not T.chain_ok(sb, adj32, N, K)
