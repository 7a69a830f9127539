"""CPU: the vectorised TU-format reader and the edge-list batch feed (graph_pooling_b200/data.py) reproduce the
reference's loader + sampler semantics (load_data.py:7-109, graph_sampler.py:26-37,97-109).

  * against tests/golden/dataset_enzymes.npz, which was produced THROUGH the reference's own loader;
  * against the reference loader itself on a synthetic dataset with the awkward cases (isolated node, a graph
    label 0, node attributes, a graph above max_nodes) -- only where /root/reference exists (build container)."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import HERE, load_enzymes

REF = '/root/reference'


def fixture_graphset():
    from graph_pooling_b200.data import GraphSet
    z = np.load(os.path.join(HERE, 'golden', 'dataset_enzymes.npz'))
    return GraphSet(z['n'], z['glabel'].astype(np.int64) - int(z['glabel'].min()), z['nlabel'], z['eptr'],
                    z['edges'], int(z['num_node_labels']))


def test_batch_feed_matches_padded_arrays():
    x, adj, nb, label = load_enzymes()
    gs = fixture_graphset().to('cpu')
    idx = np.array([5, 0, 17, 596, 300, 42, 17, 5, 5])        # with repeats: sampling with replacement / oversampled folds
    bx, badj, bnb, bl = gs.batch(idx, 100, adj_dtype=torch.float32)
    assert np.array_equal(bx.numpy(), x[idx]) and np.array_equal(badj.numpy(), adj[idx])
    assert np.array_equal(bnb.numpy(), nb[idx]) and np.array_equal(bl.numpy(), label[idx])
    u8 = gs.batch(idx, 100)[1]
    assert u8.dtype == torch.uint8 and np.array_equal(u8.numpy().astype(np.float32), adj[idx])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'data', 'ENZYMES')), reason='reference data not present')
def test_reader_matches_reference_loader_on_enzymes():
    from graph_pooling_b200.data import read_tu_dataset
    z = np.load(os.path.join(HERE, 'golden', 'dataset_enzymes.npz'))
    gs = read_tu_dataset(os.path.join(REF, 'data'), 'ENZYMES', max_nodes=100)
    assert np.array_equal(gs.n, z['n']) and np.array_equal(gs.label, z['glabel'])
    assert np.array_equal(gs.nlabel, z['nlabel']) and gs.num_node_labels == int(z['num_node_labels'])
    assert np.array_equal(gs.eptr, z['eptr'])
    for g in range(len(gs)):                              # same edge SET per graph (networkx's edge order is its own)
        a = {tuple(r) for r in gs.edges[gs.eptr[g]:gs.eptr[g + 1]]}
        b = {tuple(r) for r in z['edges'][z['eptr'][g]:z['eptr'][g + 1]].astype(np.int64)}
        assert a == b, g
    assert gs.attrs is not None and gs.attrs.shape == (int(gs.n.sum()), 18)


def _write_tu(root, name):
    d = os.path.join(root, name)
    os.makedirs(d)
    # graph 1: nodes 1-4 (node 4 isolated); graph 2: nodes 5-7; graph 3: nodes 8-13 (a 6-cycle, above max_nodes=5)
    gi = [1] * 4 + [2] * 3 + [3] * 6
    edges = [(2, 1), (1, 2), (3, 2), (2, 3), (6, 7), (7, 6), (5, 6), (6, 5), (5, 7), (7, 5)]
    cyc = list(range(8, 14))
    for i in range(6):
        edges += [(cyc[i], cyc[(i + 1) % 6]), (cyc[(i + 1) % 6], cyc[i])]
    open(os.path.join(d, name + '_graph_indicator.txt'), 'w').write('\n'.join(map(str, gi)) + '\n')
    open(os.path.join(d, name + '_A.txt'), 'w').write('\n'.join('%d, %d' % e for e in edges) + '\n')
    open(os.path.join(d, name + '_graph_labels.txt'), 'w').write('0\n1\n1\n')
    open(os.path.join(d, name + '_node_labels.txt'), 'w').write('\n'.join(str(1 + (i % 3)) for i in range(13)) + '\n')
    open(os.path.join(d, name + '_node_attributes.txt'), 'w').write(
        '\n'.join('%.1f, %.2f' % (i, -i / 2) for i in range(13)) + '\n')


def test_reader_awkward_cases(tmp_path):
    from graph_pooling_b200.data import read_tu_dataset
    _write_tu(str(tmp_path), 'TOY')
    gs = read_tu_dataset(str(tmp_path), 'TOY', max_nodes=5)
    assert gs.n.tolist() == [3, 3]                        # isolated node 4 dropped, graph 3 dropped
    assert gs.label.tolist() == [0, 1]                    # a 0 label exists -> labels kept as they are
    # node order = first appearance in the edge list: graph 1 -> [2, 1, 3], graph 2 -> [6, 7, 5]
    assert gs.nlabel.tolist() == [1, 0, 2, 2, 0, 1]
    assert np.allclose(gs.attrs[:, 0], [1, 0, 2, 5, 6, 4])
    e0 = {tuple(r) for r in gs.edges[gs.eptr[0]:gs.eptr[1]]}
    e1 = {tuple(r) for r in gs.edges[gs.eptr[1]:gs.eptr[2]]}
    assert e0 == {(0, 1), (0, 2)} and e1 == {(0, 1), (0, 2), (1, 2)}
    if os.path.isfile(os.path.join(REF, 'load_data.py')):  # and the same through the reference's own loader
        sys.path.insert(0, os.path.dirname(HERE))
        from graph_pooling_b200 import shim
        shim.install_networkx_compat()
        sys.path.insert(0, REF)
        import load_data
        graphs = load_data.read_graphfile(str(tmp_path), 'TOY', max_nodes=5)
        assert [g.number_of_nodes() for g in graphs] == gs.n.tolist()
        assert [int(g.graph['label']) for g in graphs] == gs.label.tolist()
        off = 0
        for gi_, g in enumerate(graphs):
            assert [int(np.argmax(g.node[u]['label'])) for u in range(g.number_of_nodes())] == \
                gs.nlabel[off:off + gs.n[gi_]].tolist()
            assert np.allclose([g.node[u]['feat'] for u in range(g.number_of_nodes())], gs.attrs[off:off + gs.n[gi_]])
            assert {(min(u, v), max(u, v)) for u, v in g.edges()} == \
                {tuple(r) for r in gs.edges[gs.eptr[gi_]:gs.eptr[gi_ + 1]]}
            off += gs.n[gi_]


@pytest.mark.parametrize('rows,N,threads', [(37, 100, 1), (64, 2048, 4), (5, 13, 3), (200, 64, 16)])
def test_host_bit_packer_matches_numpy(rows, N, threads):
    """gp_host_pack_adj_bits (host code of the feed, no GPU needed): bit c & 7 of byte c >> 3, zero padding to the row
    stride, and the 'entry outside {0,1}' report."""
    import ctypes as C
    from graph_pooling_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(rows + N)
    a = (rs.rand(rows, N) < 0.3).astype(np.float32)
    ldb = (N + 7) // 8 + 3
    out = np.full((rows, ldb), 0xAB, np.uint8)
    bad = C.c_int(-1)
    rc = lib.gp_host_pack_adj_bits(a.ctypes.data, rows, N, out.ctypes.data, ldb, threads, C.addressof(bad))
    assert rc == 0 and bad.value == 0
    ref = np.packbits(a.astype(np.uint8), axis=1, bitorder='little')
    assert np.array_equal(out[:, :ref.shape[1]], ref) and not out[:, ref.shape[1]:].any()
    a[rows // 2, N // 3] = 0.5
    lib.gp_host_pack_adj_bits(a.ctypes.data, rows, N, out.ctypes.data, ldb, threads, C.addressof(bad))
    assert bad.value == 1
