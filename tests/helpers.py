"""Shared helpers for the parity tests (golden loading, error metrics, seeded batches)."""
import glob
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, 'golden', '*.npz'))
                if not os.path.basename(p).startswith('dataset_'))


def load_golden(name):
    z = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    return {k: z[k] for k in z.files}


def golden_model(g, mod, dtype=torch.float32, device='cpu'):
    """Build `mod`'s encoder for golden case `g` and load the reference's state dict."""
    B, N, D, H, E, C, L = [int(v) for v in g['meta']]
    if str(g['kind']) == 'soft':
        m = mod.SoftPoolingGcnEncoder(N, D, H, E, C, L, H, assign_ratio=float(g['ratio']), num_pooling=1,
                                      bn=True, linkpred=True, assign_input_dim=D)
    elif str(g['kind']) == 's2s':
        m = mod.GcnSet2SetEncoder(D, H, E, C, L, bn=True)
    else:
        m = mod.GcnEncoderGraph(D, H, E, C, L, bn=bool(g['bn']) if 'bn' in g else True,
                                concat=bool(g['concat']) if 'concat' in g else True)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')}
    m.load_state_dict(sd, strict=True)
    return m.to(device=device, dtype=dtype)


def golden_inputs(g, dtype=torch.float32, device='cpu'):
    x = torch.from_numpy(g['x']).to(device=device, dtype=dtype)
    adj = torch.from_numpy(g['adj_u8'].astype(np.float32)).to(device=device, dtype=dtype)
    label = torch.from_numpy(g['label']).to(device)
    return x, adj, g['nb'], label


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def synth_batch(seed, B, N, D, n_min, n_max, C, density=0.1, symmetric=True, weighted=False):
    rs = np.random.RandomState(seed)
    nb = rs.randint(n_min, n_max + 1, size=B).astype(np.int32)
    adj = np.zeros((B, N, N), np.float32)
    x = np.zeros((B, N, D), np.float32)
    for b in range(B):
        n = int(nb[b])
        a = (rs.rand(n, n) < density).astype(np.float32)
        if weighted:
            a = a * rs.rand(n, n).astype(np.float32)
        if symmetric:
            u = np.triu(a, 1)
            a = u + u.T
        adj[b, :n, :n] = a
        x[b, :n] = rs.randn(n, D).astype(np.float32)
    label = rs.randint(0, C, size=B).astype(np.int64)
    return x, adj, nb, label


def load_enzymes(max_nodes=100):
    """tests/golden/dataset_enzymes.npz (made by make_enzymes_fixture.py through the reference's loader) -> padded
    arrays exactly as graph_sampler.py:97-109 + train.py:477-481 feed them: dense {0,1} adjacency, one-hot
    node-label features, zero padding.  Returns x [G,N,D], adj [G,N,N], nb [G], label [G] (0-based)."""
    z = np.load(os.path.join(HERE, 'golden', 'dataset_enzymes.npz'))
    n, eptr, edges, nlabel = z['n'].astype(np.int64), z['eptr'], z['edges'].astype(np.int64), z['nlabel']
    G, D = len(n), int(z['num_node_labels'])
    adj = np.zeros((G, max_nodes, max_nodes), np.float32)
    x = np.zeros((G, max_nodes, D), np.float32)
    off = 0
    for g in range(G):
        e = edges[eptr[g]:eptr[g + 1]]
        adj[g, e[:, 0], e[:, 1]] = 1.0
        adj[g, e[:, 1], e[:, 0]] = 1.0
        x[g, np.arange(n[g]), nlabel[off:off + n[g]]] = 1.0
        off += n[g]
    label = z['glabel'].astype(np.int64)
    return x, adj, n.astype(np.int32), label - label.min()
