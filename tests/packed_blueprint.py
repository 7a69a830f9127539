"""Blueprint (test infrastructure, CPU, float64) of the PACKED small-graph schedule implemented by csrc/packed_*.cu.

It restates, phase by phase and graph by graph, exactly what the packed kernels compute for a DiffPool step
(SoftPoolingGcnEncoder, num_pooling = 1, concat, BatchNorm) -- only the real n_b rows of every graph exist, pad rows
are represented by their per-layer constant normalize(b) and a count per node index -- with hand-written backward
formulas (no autograd).  tests/test_packed_blueprint_cpu.py checks it against the oracle's autograd results, so the
data flow and every formula the CUDA kernels transcribe are pinned on the CPU.

Phases (one kernel launch each; a launch boundary is the grid-wide dependency BatchNorm needs):
  F(l), l = 0..L-1   level-0 layer l of the embedding and assignment GCN in lock-step; the last one also does the
                     readout, assignment softmax, pooling, link loss and layer 0 of the post-pool GCN
  P(l), l = 1..L-1   post-pool GCN layer l (the last one: readout)
  head               prediction MLP + cross entropy, forward and backward
  BP(l), l = L-1..0  post-pool GCN backward
  BPOOL              pooling / softmax / assign_pred / link-loss backward + level-0 last layer backward
  B(l), l = L-2..0   level-0 layer backward
"""
import numpy as np
import torch

EPS_NORM, EPS_BN, EPS_LINK = 1e-12, 1e-5, 1e-7


def _norm_rows(v):
    r = v.norm(dim=-1, keepdim=True).clamp_min(EPS_NORM)
    return v / r, r


def _norm_bwd(dy, y, r):
    # rows whose norm was clamped (||V|| < eps): Y = V / eps, dV = dY / eps
    proj = (y * dy).sum(-1, keepdim=True)
    return torch.where(r > EPS_NORM, (dy - y * proj) / r, dy / EPS_NORM)


class Stack:
    """Weights of one GCN stack: list of (W [in,out], b [out] or None)."""

    def __init__(self, layers):
        self.W = [w for w, _ in layers]
        self.b = [b for _, b in layers]
        self.L = len(layers)

    def pad_y(self, l):
        b = self.b[l]
        if b is None:
            return torch.zeros(self.W[l].shape[1], dtype=self.W[l].dtype)
        return b / b.norm().clamp_min(EPS_NORM)


def stack_forward(st, A, X, nb, cnt_pad, B):
    """A: list of [n,n]; X: list of [n,din].  Returns per-layer saved state and the concat Z per graph.
    BatchNorm per node index over (B graphs, d features), pad rows (cnt_pad[n] graphs at index n) included
    analytically."""
    L = st.L
    G = len(A)
    Hin = X
    Ys, Rn, stats, Hs = [], [], [], []
    for l in range(L):
        Y, R = [], []
        for g in range(G):
            y, r = _norm_rows(A[g] @ Hin[g] @ st.W[l] + (0 if st.b[l] is None else st.b[l]))
            Y.append(y)
            R.append(r)
        Ys.append(Y)
        Rn.append(R)
        if l == L - 1:
            break
        d = st.W[l].shape[1]
        Nmax = len(cnt_pad)
        s1 = torch.zeros(Nmax, dtype=torch.float64)
        s2 = torch.zeros(Nmax, dtype=torch.float64)
        for g in range(G):
            rl = torch.relu(Y[g])
            s1[:nb[g]] += rl.sum(1)
            s2[:nb[g]] += (rl * rl).sum(1)
        rp = torch.relu(st.pad_y(l))
        s1 += cnt_pad * rp.sum()
        s2 += cnt_pad * (rp * rp).sum()
        mean = s1 / (B * d)
        var = s2 / (B * d) - mean * mean
        istd = 1.0 / torch.sqrt(var + EPS_BN)
        stats.append((mean, istd))
        H = [(torch.relu(Y[g]) - mean[:nb[g], None]) * istd[:nb[g], None] for g in range(G)]
        Hs.append(H)
        Hin = H
    Z = [torch.cat([Hs[l][g] for l in range(L - 1)] + [Ys[L - 1][g]], dim=1) for g in range(G)]
    return dict(Y=Ys, R=Rn, stats=stats, H=Hs, Z=Z, X=X)


def stack_backward(st, A, sv, gz, nb, cnt_pad, B, need_dx, need_da):
    """gz: list of [n, F] upstream gradients of the (masked) concat.  Returns (dW list, db list, dX list, dA list)."""
    L = st.L
    G = len(A)
    widths = [w.shape[1] for w in st.W]
    offs = np.concatenate([[0], np.cumsum(widths)])
    dW = [torch.zeros_like(w) for w in st.W]
    db = [None if b is None else torch.zeros_like(b) for b in st.b]
    dA = [torch.zeros_like(a) for a in A] if need_da else None
    dxn = [None] * G
    for l in reversed(range(L)):
        gl = [gz[g][:, offs[l]:offs[l + 1]] + (0 if dxn[g] is None else dxn[g]) for g in range(G)]
        Hin = sv['X'] if l == 0 else sv['H'][l - 1]
        if l < L - 1:
            mean, istd = sv['stats'][l]
            d = widths[l]
            Nmax = len(cnt_pad)
            m1 = torch.zeros(Nmax, dtype=torch.float64)
            m2 = torch.zeros(Nmax, dtype=torch.float64)
            for g in range(G):
                m1[:nb[g]] += gl[g].sum(1)
                m2[:nb[g]] += (gl[g] * sv['H'][l][g]).sum(1)
            m1 /= (B * d)
            m2 /= (B * d)
            dY = []
            for g in range(G):
                n = nb[g]
                dR = (gl[g] - m1[:n, None] - sv['H'][l][g] * m2[:n, None]) * istd[:n, None]
                dY.append(dR * (sv['Y'][l][g] > 0))
            # pad rows: upstream 0, but the batch means reach them; one vector per node index, cnt_pad[n] copies
            yp = st.pad_y(l)
            if st.b[l] is not None:
                hp = (torch.relu(yp)[None, :] - mean[:, None]) * istd[:, None]          # [Nmax, d]
                dRp = (-m1[:, None] - hp * m2[:, None]) * istd[:, None]
                dYp = dRp * (yp > 0)[None, :]
                rp = st.b[l].norm().clamp_min(EPS_NORM)
                dVp = _norm_bwd(dYp, yp[None, :].expand_as(dYp), rp.expand(len(cnt_pad), 1))
                db[l] += (cnt_pad[:, None] * dVp).sum(0)
        else:
            dY = gl
        for g in range(G):
            dV = _norm_bwd(dY[g], sv['Y'][l][g], sv['R'][l][g])
            U = A[g] @ Hin[g]
            dW[l] += U.t() @ dV
            if db[l] is not None:
                db[l] += dV.sum(0)
            if l > 0 or need_dx or need_da:
                dU = dV @ st.W[l].t()
                dxn[g] = A[g].t() @ dU
                if need_da:
                    dA[g] += dU @ Hin[g].t()
    return dW, db, dxn, dA


def diffpool_step(params, x, adj, nb, label, assign_x=None):
    """params: dict with Stack 'emb', 'assign', 'post', (Wp [K,Fa], bp), and MLP list [(W [out,in], b)].
    x: [B,N,D] / adj: [B,N,N] padded float64 tensors; nb: int array.  Returns dict(ypred, loss, link, S, grads...)."""
    B, N = x.shape[0], x.shape[1]
    nb = [int(v) for v in nb]
    cnt_pad = torch.tensor([sum(1 for v in nb if v <= n) for n in range(N)], dtype=torch.float64)
    A = [adj[g, :nb[g], :nb[g]] for g in range(B)]
    X = [x[g, :nb[g]] for g in range(B)]
    XA = X if assign_x is None else [assign_x[g, :nb[g]] for g in range(B)]
    emb, asg, post = params['emb'], params['assign'], params['post']
    Wp, bp = params['assign_pred']
    # ---- forward, level 0
    se = stack_forward(emb, A, X, nb, cnt_pad, B)
    sa = stack_forward(asg, A, XA, nb, cnt_pad, B)
    F = se['Z'][0].shape[1]
    out0 = torch.zeros(B, F, dtype=torch.float64)
    arg0 = torch.zeros(B, F, dtype=torch.long)
    for g in range(B):
        z = se['Z'][g]
        m, a = z.max(0)
        if nb[g] < N:                                   # zeroed pad rows take part in the max (encoders.py:1257)
            pad_wins = m < 0
            m = torch.where(pad_wins, torch.zeros_like(m), m)
            a = torch.where(pad_wins, torch.full_like(a, -1), a)
        out0[g], arg0[g] = m, a
    S, Xp, Ap = [], [], []
    link_sum = torch.zeros((), dtype=torch.float64)
    for g in range(B):
        t = sa['Z'][g] @ Wp.t() + (0 if bp is None else bp)
        s = torch.softmax(t, dim=-1)
        S.append(s)
        Xp.append(s.t() @ se['Z'][g])
        Ap.append(s.t() @ A[g] @ s)
        P = torch.clamp(s @ s.t(), max=1.0)
        link_sum += (-A[g] * torch.log(P + EPS_LINK) - (1 - A[g]) * torch.log(1 - P + EPS_LINK)).sum()
    entries = float(sum(v * v for v in nb))
    link = link_sum / entries
    K = Wp.shape[0]
    nbk = [K] * B
    zero_cnt = torch.zeros(K, dtype=torch.float64)
    sp = stack_forward(post, Ap, Xp, nbk, zero_cnt, B)
    out1 = torch.stack([sp['Z'][g].max(0)[0] for g in range(B)])
    arg1 = torch.stack([sp['Z'][g].max(0)[1] for g in range(B)])
    # ---- head
    h = torch.cat([out0, out1], dim=1)
    acts = [h]
    mlp = params['pred']
    for i, (w, b) in enumerate(mlp):
        h = h @ w.t() + b
        if i < len(mlp) - 1:
            h = torch.relu(h)
        acts.append(h)
    ypred = h
    logp = torch.log_softmax(ypred, dim=1)
    ce = -logp[torch.arange(B), label].mean()
    loss = ce + link
    # head backward
    gq = torch.softmax(ypred, dim=1)
    gq[torch.arange(B), label] -= 1.0
    gq /= B
    gmlp = [None] * len(mlp)
    for i in reversed(range(len(mlp))):
        w, b = mlp[i]
        if i < len(mlp) - 1:
            gq = gq * (acts[i + 1] > 0)
        gmlp[i] = (gq.t() @ acts[i], gq.sum(0))
        gq = gq @ w
    dout = gq                                               # [B, 2F]
    # ---- post stack backward
    gz1 = []
    for g in range(B):
        gz = torch.zeros(K, F, dtype=torch.float64)
        gz[arg1[g], torch.arange(F)] = dout[g, F:]
        gz1.append(gz)
    dWq, dbq, dXp, dAp = stack_backward(post, Ap, sp, gz1, nbk, zero_cnt, B, True, True)
    # ---- pooling / softmax / assign_pred / link backward
    gze, gza = [], []
    dWp = torch.zeros_like(Wp)
    dbp = None if bp is None else torch.zeros_like(bp)
    for g in range(B):
        s, z, a = S[g], se['Z'][g], A[g]
        dz = s @ dXp[g]
        ds = z @ dXp[g].t() + a @ s @ dAp[g].t() + a.t() @ s @ dAp[g]
        Praw = s @ s.t()
        P = torch.clamp(Praw, max=1.0)
        Gm = (-a / (P + EPS_LINK) + (1 - a) / (1 - P + EPS_LINK)) * (Praw <= 1.0) / entries
        ds = ds + (Gm + Gm.t()) @ s
        dt = s * (ds - (ds * s).sum(-1, keepdim=True))
        dWp += dt.t() @ sa['Z'][g]
        if dbp is not None:
            dbp += dt.sum(0)
        gza.append(dt @ Wp)
        gz = dz.clone()
        for f in range(F):
            if arg0[g, f] >= 0:
                gz[arg0[g, f], f] += dout[g, f]
        gze.append(gz)
    dWe, dbe, _, _ = stack_backward(emb, A, se, gze, nb, cnt_pad, B, False, False)
    dWa, dba, _, _ = stack_backward(asg, A, sa, gza, nb, cnt_pad, B, False, False)
    return dict(ypred=ypred, loss=loss, link=link, S=S, emb=(dWe, dbe), assign=(dWa, dba), post=(dWq, dbq),
                assign_pred=(dWp, dbp), pred=gmlp)
