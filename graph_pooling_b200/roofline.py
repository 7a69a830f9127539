"""Algorithmic FLOP / byte model of one DiffPool train step; the single source for
bench.py's roofline numbers.  "Useful" work only: real n_b x n_b blocks, reference association
(A.X).W, fwd + bwd, no dA at level 0; BN / normalize / softmax count as 0 flops."""
import numpy as np


def _stack(n, dims):
    """sum over layers of 2 n^2 in + 2 n in out"""
    return sum(2.0 * n * n * i + 2.0 * n * i * o for i, o in dims)


def _layer_dims(din, H, E, L):
    return [(din, H)] + [(H, H)] * (L - 2) + [(H, E)]


def graph_flops(n, cfg):
    """(fwd, bwd) useful flops for one graph with n real nodes."""
    D, H, E_, L = cfg['D'], cfg['H'], cfg['E'], cfg['L']
    F = H * (L - 1) + E_
    emb = _layer_dims(D, H, E_, L)
    fwd = _stack(n, emb)
    bwd = sum(4.0 * n * i * o + (0.0 if k == 0 else 2.0 * n * n * i) for k, (i, o) in enumerate(emb))
    if cfg['kind'] != 'soft':
        return fwd, bwd
    K = int(cfg['N'] * cfg['ratio'])
    cur_n, cur_in, first = float(n), D, True
    for lvl in range(cfg['P']):
        Fa = H * (L - 1) + K
        asg = _layer_dims(cur_in, H, K, L)
        post = _layer_dims(F, H, E_, L)
        fwd += _stack(cur_n, asg) + 2 * cur_n * Fa * K + 2 * cur_n * K * F + 2 * K * cur_n ** 2 + 2 * K * K * cur_n
        fwd += _stack(K, post)
        bwd += sum(4.0 * cur_n * i * o + (0.0 if (k == 0 and first) else 2.0 * cur_n ** 2 * i)
                   for k, (i, o) in enumerate(asg))
        bwd += 4 * cur_n * Fa * K + 4 * cur_n * K * F + 2 * K * cur_n ** 2 + 4 * K * K * cur_n
        bwd += sum(4.0 * K * i * o + 2.0 * K * K * i + 2.0 * K * K * i for i, o in post)   # dX' and dA'
        if lvl == 0:                                                         # link loss on level-0 S
            fwd += 2 * cur_n ** 2 * K
            bwd += 4 * cur_n ** 2 * K
        cur_n, cur_in, first = float(K), F, False
        K = int(K * cfg['ratio'])
    return fwd, bwd


def step_flops(nb, cfg):
    f = b = 0.0
    for n in np.asarray(nb, dtype=np.float64):
        ff, bb = graph_flops(n, cfg)
        f += ff
        b += bb
    return f, b


def step_bytes(nb, cfg, elt=4):
    """Compulsory HBM traffic of one train step (A read once fwd + once bwd, inputs, saved activations
    written once and read once)."""
    D, H, E_, L = cfg['D'], cfg['H'], cfg['E'], cfg['L']
    F = H * (L - 1) + E_
    K = int(cfg['N'] * cfg['ratio']) if cfg['kind'] == 'soft' else 0
    Fa = H * (L - 1) + K if K else 0
    tot = 0.0
    for n in np.asarray(nb, dtype=np.float64):
        tot += 2 * n * n * elt + 2 * n * D * elt + 2 * n * (F + Fa + K) * elt
    return tot


def ax_kernel_work(nb, din, elt=4):
    """Algorithmic work of ONE launch of the dominant contraction U = A.X over the batch:
    flops = sum 2 n_b^2 din ; bytes = sum (n_b^2 + 2 n_b din) * elt."""
    nbf = np.asarray(nb, dtype=np.float64)
    return float(np.sum(2 * nbf * nbf * din)), float(np.sum((nbf * nbf + 2 * nbf * din) * elt))
