"""Seeded synthetic padded graph batches for the BASELINE.json configs.

Contract of the reference's feed (graph_sampler.py:97-109): dense fp32 adjacency, symmetric {0,1},
zero diagonal, zero rows/columns beyond each graph's node count; features zero on pad rows.
"""
import numpy as np
import torch

# name -> dict(kind, B, N, D, H, E, C, L, ratio, P, n_min, n_max, density | mean_degree)
WORKLOADS = {
    # configs[3]: synthetic padded batch 256 graphs x 2048 nodes, hidden 128, assign-ratio 0.25
    'cfg4_diffpool_256x2048': dict(kind='soft', B=256, N=2048, D=128, H=128, E=128, C=2, L=3, ratio=0.25, P=1,
                                   n_min=2048, n_max=2048, density=0.01),
    # configs[0]: DiffPool on ENZYMES shapes (batch 20, hidden/output 30, assign-ratio 0.1, num_pool 1)
    'cfg1_enzymes_like': dict(kind='soft', B=20, N=100, D=3, H=30, E=30, C=6, L=3, ratio=0.1, P=1,
                              n_min=2, n_max=100, mean_degree=3.9, enzymes_hist=True),
    # configs[1]: base GCN on DD shapes (max_nodes 1000, 3 GraphConv layers, hidden 20)
    'cfg2_dd_base': dict(kind='base', B=20, N=1000, D=89, H=20, E=20, C=2, L=3, ratio=0.0, P=0,
                         n_min=30, n_max=1000, mean_degree=5.0, dd_hist=True),
    # configs[2]: DiffPool on DD shapes, num_pool 2, assign-ratio 0.25, hidden 64
    'cfg3_dd_diffpool_p2': dict(kind='soft', B=20, N=1000, D=89, H=64, E=64, C=2, L=3, ratio=0.25, P=2,
                                n_min=30, n_max=1000, mean_degree=5.0, dd_hist=True),
    # configs[4]: ragged batch, n_b ~ U{50..5000}
    'cfg5_ragged_64x5000': dict(kind='soft', B=64, N=5000, D=128, H=128, E=128, C=2, L=3, ratio=0.25, P=1,
                                n_min=50, n_max=5000, mean_degree=8.0),
    # configs[0] on the REAL bundled ENZYMES graphs (tests/golden/dataset_enzymes.npz, produced through the reference's
    # own loader): batches of real graphs assembled on the device by data.GraphSet
    'cfg1_enzymes_real': dict(kind='soft', B=20, N=100, D=3, H=30, E=30, C=6, L=3, ratio=0.1, P=1,
                              n_min=2, n_max=100, fixture='dataset_enzymes.npz'),
    # small smoke-sized DiffPool
    'tiny': dict(kind='soft', B=8, N=64, D=8, H=16, E=16, C=3, L=3, ratio=0.25, P=1, n_min=4, n_max=64,
                 density=0.1),
}


def node_counts(cfg, B, rs):
    n_min, n_max = cfg['n_min'], cfg['n_max']
    if cfg.get('enzymes_hist'):
        # ENZYMES node-count histogram: min 2, mean ~32, p95 ~54, max 100; lognormal fit
        n = np.exp(rs.normal(np.log(30.0), 0.45, size=B))
    elif cfg.get('dd_hist'):
        # DD node counts <= 1000: mean ~269, median ~241, p95 ~589
        n = np.exp(rs.normal(np.log(235.0), 0.55, size=B))
    else:
        n = rs.randint(n_min, n_max + 1, size=B).astype(np.float64)
    return np.clip(np.round(n), n_min, n_max).astype(np.int32)


def make_batch(name, seed=0, device='cuda', B=None):
    """Returns dict(x, adj, nb (numpy int32), label, cfg) with tensors on `device` (generated there)."""
    cfg = dict(WORKLOADS[name])
    if B is not None:
        cfg['B'] = B
    B, N, D, C = cfg['B'], cfg['N'], cfg['D'], cfg['C']
    rs = np.random.RandomState(seed)
    if cfg.get('fixture'):
        import os
        from .data import GraphSet
        z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden',
                                 cfg['fixture']))
        gs = GraphSet(z['n'], z['glabel'].astype(np.int64) - int(z['glabel'].min()), z['nlabel'], z['eptr'],
                      z['edges'], int(z['num_node_labels'])).to(device)
        idx = rs.choice(len(gs), size=B, replace=B > len(gs))
        x, adj, nbd, label = gs.batch(idx, N, adj_dtype=torch.float32)
        return dict(x=x.contiguous(), adj=adj, nb=nbd.cpu().numpy().astype(np.int32), label=label, cfg=cfg)
    nb = node_counts(cfg, B, rs)
    g = torch.Generator(device=device).manual_seed(seed)
    nbt = torch.as_tensor(nb.astype(np.int64), device=device)
    idx = torch.arange(N, device=device)
    real = idx[None, :] < nbt[:, None]                                       # [B,N]
    adj = torch.empty(B, N, N, device=device, dtype=torch.float32)
    for b in range(B):                                                       # per graph: bounded temporaries
        n = int(nb[b])
        p = cfg['density'] if 'density' in cfg else min(1.0, cfg['mean_degree'] / max(n - 1, 1))
        u = (torch.rand(N, N, device=device, generator=g) < p)
        u = torch.triu(u, diagonal=1) & real[b][None, :] & real[b][:, None]
        adj[b] = (u | u.t()).float()
    if cfg['D'] <= 100 and cfg.get('dd_hist') or cfg.get('enzymes_hist'):
        lab = torch.randint(0, D, (B, N), device=device, generator=g)
        x = torch.nn.functional.one_hot(lab, D).float()
    else:
        x = torch.randn(B, N, D, device=device, generator=g)
    x = x * real[:, :, None].float()
    label = torch.randint(0, C, (B,), device=device, generator=g)
    return dict(x=x.contiguous(), adj=adj, nb=nb, label=label, cfg=cfg)


def build_model(mod, cfg):
    """Construct `mod`'s encoder (mod = graph_pooling_b200.encoders or the oracle) for a workload."""
    if cfg['kind'] == 'soft':
        return mod.SoftPoolingGcnEncoder(cfg['N'], cfg['D'], cfg['H'], cfg['E'], cfg['C'], cfg['L'], cfg['H'],
                                         assign_ratio=cfg['ratio'], num_pooling=cfg['P'], bn=True, linkpred=True,
                                         assign_input_dim=cfg['D'])
    return mod.GcnEncoderGraph(cfg['D'], cfg['H'], cfg['E'], cfg['C'], cfg['L'], bn=True)
