"""PACKED schedule for ENZYMES-sized graphs (csrc/packed.cu; include/gp_b200.h "PACKED schedule").

DiffPool with one pooling level (SoftPoolingGcnEncoder, encoders.py:1231-1300) on graphs of at most 128 nodes: only
the real n_b rows of every graph exist, one launch per phase for the whole batch (15 phases + the prediction MLP),
BatchNorm statistics per node index carried from phase to phase as sums.  fp32 FFMA throughout: this IS the fp32
parity mode for small graphs (tests/test_gpu_packed.py against the oracle), selected automatically by
encoders._EncoderFn when `supported()` says so (GP_NO_PACKED=1 switches it off).

Nothing here computes: every tensor op is a call into the C ABI.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import engine as E
from ._lib import (GpPkAdj, GpPkConcat, GpPkGrad, GpPkLayerBwdArgs, GpPkLayerFwdArgs, GpPkPoolArgs, GpPkSrc,
                   GpPkTiling, PK_ELL, PK_MAX_LAYERS, call)

MAX_N = 128           # kMaxN in packed.cu
WINDOW = int(os.environ.get('GP_PK_WINDOW', 96))           # packed rows per window (the graphs whose first row falls into it); gp_pk_prepare splits every
                      # window into runs of whole graphs of at most max(N, WINDOW) rows: the unit a CTA works on
POST_ROWS = int(os.environ.get('GP_PK_POST_ROWS', 64))        # rows per run at the pooled level (whole graphs of K rows)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def supported(plan, x, adj, assign_x, params):
    """The packed schedule covers: soft-assign, one pooling level, concat, BatchNorm, no dropout, no add_self, fp32
    tensors, N <= 128 and every width <= 128."""
    if os.environ.get('GP_NO_PACKED'):
        return False
    if not (plan.soft and plan.num_pooling == 1 and plan.concat and plan.bn and plan.bn_post and not plan.add_self):
        return False
    if plan.precision != E.F32 or not torch.is_tensor(adj) or adj.dtype != torch.float32:
        return False
    if any(plan.emb_drop) or any(any(d) for d in plan.post_drop):
        return False
    B, N, D = x.shape
    K = plan.assign_dims[0]
    if N > MAX_N or K > MAX_N or K < 1 or D > MAX_N or assign_x.shape[2] > MAX_N:
        return False
    for pairs in (plan.emb, plan.assign[0], plan.post[0]):
        if len(pairs) > PK_MAX_LAYERS or len(pairs) < 2:
            return False
        for iw, _ in pairs:
            w = params[iw]
            if w.shape[0] > 512 or w.shape[1] > MAX_N or not w.is_contiguous():
                return False
    if plan.F > 256 or (B * N * N) >= (1 << 31):
        return False
    return True


def _src(y_ptr, ld, d, sums=None, bias=None):
    return GpPkSrc(y_ptr, ld, d, sums, bias)


def _r4(v):
    return (int(v) + 3) & ~3


class _Stack:
    """One GCN stack on packed rows: weights, per-layer Y / rnorm / BatchNorm sums."""

    def __init__(self, ws, rows, weights, biases, in_src, dbl, dbl_off, Nn):
        self.W, self.b, self.L = weights, biases, len(weights)
        self.d = [int(w.shape[1]) for w in weights]
        self.din = [int(w.shape[0]) for w in weights]
        self.F = sum(self.d)
        self.in_src = in_src
        self.Y = [ws.f(rows, d) for d in self.d]
        self.rn = [ws.f(rows) for _ in self.d]
        # sums[l] / msums[l] for l < L-1: [2*Nn] doubles each inside the zeroed blob
        self.sums, self.msums = [], []
        base = dbl.data_ptr()
        for l in range(self.L - 1):
            self.sums.append(base + 8 * (dbl_off + (2 * l) * 2 * Nn))
            self.msums.append(base + 8 * (dbl_off + (2 * l + 1) * 2 * Nn))
        self.sums.append(None)
        self.msums.append(None)

    @staticmethod
    def doubles(L, Nn):
        return 2 * (L - 1) * 2 * Nn

    def src_in(self, l):
        if l == 0:
            return self.in_src
        return _src(self.Y[l - 1].data_ptr(), self.d[l - 1], self.d[l - 1], self.sums[l - 1], _p(self.b[l - 1]))

    def src_out(self, l):
        return _src(self.Y[l].data_ptr(), self.d[l], self.d[l], self.sums[l], _p(self.b[l]))

    def concat(self):
        c = GpPkConcat()
        c.L, c.F = self.L, self.F
        for l in range(self.L):
            c.slot[l] = self.src_out(l)
        return c

    def offs(self):
        return [int(v) for v in np.concatenate([[0], np.cumsum(self.d)])]


def _layer_fwd(tl, adj, cnt_pad, Nn, stacks, l):
    a = GpPkLayerFwdArgs()
    a.tl, a.adj, a.cnt_pad, a.N = tl, adj, cnt_pad, Nn
    act = [s for s in stacks if l < s.L]
    a.ns = len(act)
    for i, s in enumerate(act):
        f = a.s[i]
        f.inp = s.src_in(l)
        f.W, f.b, f.dout = s.W[l].data_ptr(), _p(s.b[l]), s.d[l]
        f.y, f.rnorm, f.sums_out = s.Y[l].data_ptr(), s.rn[l].data_ptr(), s.sums[l]
    call('gp_pk_layer_fwd', C.byref(a), _stream())


def forward(plan, x, adj, assign_x, params, wb):
    """Returns (ypred, S [B,N,K], tape)."""
    st = _stream()
    dev = x.device
    ws = E.Workspace(dev)
    B, N, D = x.shape
    K = plan.assign_dims[0]
    nb = plan.nb_dev
    conv = lambda pairs: ([wb(params, p)[0] for p in pairs], [wb(params, p)[1] for p in pairs])
    we, be = conv(plan.emb)
    wa, ba = conv(plan.assign[0])
    wq, bq = conv(plan.post[0])
    wp, bp = wb(params, plan.assign_pred[0])
    lin = [wb(params, p) for p in plan.pred]
    Le, La, Lq = len(we), len(wa), len(wq)

    host = plan.nb_host
    rows = int(host.astype(np.int64).sum()) if host is not None else B * N
    cap = int((host.astype(np.int64) ** 2).sum()) if host is not None else B * N * N
    rows, cap = max(rows, 1), max(cap, 1)
    max_rows = max(N, WINDOW)
    nsub_max = 4 * ((rows + WINDOW - 1) // WINDOW + 1)

    # ---- one zeroed blob: BatchNorm sums (doubles) | parameter gradients (floats)
    nd_e, nd_a, nd_q = _Stack.doubles(Le, N), _Stack.doubles(La, N), _Stack.doubles(Lq, K)
    nd = nd_e + nd_a + nd_q
    shapes = [tuple(p.shape) for p in params]
    sizes = [int(np.prod(s)) for s in shapes]
    goff = [int(v) for v in np.concatenate([[0], np.cumsum(sizes)])]
    dbl = torch.empty(nd + (goff[-1] + 1) // 2 + 1, device=dev, dtype=torch.float64)
    call('gp_fill_f32', dbl.data_ptr(), C.c_longlong(2 * dbl.numel()), C.c_float(0.0), st)
    gflat = dbl[nd:].view(torch.float32)

    # ---- rows, runs, neighbour lists, packed inputs
    subs = ws.i(nsub_max, 4)                          # 16-byte aligned (fresh allocation)
    imeta = ws.i(B + 1 + 4)
    rowptr, meta = imeta[:B + 1], imeta[B + 1:]
    cnt_pad = ws.f(N)
    call('gp_pk_prepare', _p(nb), B, N, WINDOW, max_rows, rowptr.data_ptr(), cnt_pad.data_ptr(), subs.data_ptr(),
         meta.data_ptr(), st)
    Da = int(assign_x.shape[2])
    own_ax = assign_x.data_ptr() != x.data_ptr()
    info = ws.i(2, rows, 2)
    ell = ws.i(2, rows, PK_ELL, 2)
    ovf = ws.i(2, cap, 2)
    rowmeta = ws.i(rows, 2)
    xpack = ws.f(rows, _r4(D))
    axpack = ws.f(rows, _r4(Da)) if own_ax else None
    call('gp_pk_build_lists', adj.data_ptr(), _p(nb), rowptr.data_ptr(), B, N, info[0].data_ptr(), ell[0].data_ptr(),
         ovf[0].data_ptr(), info[1].data_ptr(), ell[1].data_ptr(), ovf[1].data_ptr(), meta.data_ptr() + 8,
         C.c_longlong(cap), rowmeta.data_ptr(), x.data_ptr(), D, xpack.data_ptr(), C.c_longlong(_r4(D)),
         assign_x.data_ptr() if own_ax else None, Da, _p(axpack), C.c_longlong(_r4(Da)), st)
    a_out = GpPkAdj(info[0].data_ptr(), ell[0].data_ptr(), ovf[0].data_ptr(), None, 0)
    a_in = GpPkAdj(info[1].data_ptr(), ell[1].data_ptr(), ovf[1].data_ptr(), None, 0)
    tl1 = GpPkTiling(rowptr.data_ptr(), subs.data_ptr(), meta.data_ptr() + 4, rowmeta.data_ptr(), B, N, nsub_max,
                     max_rows)
    tl2 = tl1

    # ---- level 0: embedding and assignment GCN in lock-step
    se = _Stack(ws, rows, we, be, _src(xpack.data_ptr(), _r4(D), D), dbl, 0, N)
    sa = _Stack(ws, rows, wa, ba, _src(axpack.data_ptr(), _r4(Da), Da) if own_ax else _src(xpack.data_ptr(), _r4(D), D),
                dbl, nd_e, N)
    for l in range(max(Le, La)):
        _layer_fwd(tl1, a_out, cnt_pad.data_ptr(), N, [se, sa], l)

    # ---- assignment softmax, readout, pooling
    F, Fa = se.F, sa.F
    ldo = 2 * F
    S = ws.f(B, N, K)
    xp, ap = ws.f(B, K, F), ws.f(B, K, K)
    out, arg = ws.f(B, ldo), ws.i(B, ldo)
    pa = GpPkPoolArgs()
    pa.tl, pa.adj, pa.adj_in, pa.cnt_pad, pa.nb, pa.N, pa.K = tl2, a_out, a_in, cnt_pad.data_ptr(), _p(nb), N, K
    pa.z, pa.za = se.concat(), sa.concat()
    pa.Wp, pa.bp = wp.data_ptr(), _p(bp)
    pa.S, pa.xp, pa.ap = S.data_ptr(), xp.data_ptr(), ap.data_ptr()
    pa.out, pa.arg, pa.ldo = out.data_ptr(), arg.data_ptr(), ldo
    call('gp_pk_pool_fwd', C.byref(pa), st)

    # ---- pooled level: K rows per graph, dense A'
    gpt = max(1, POST_ROWS // K)
    tlq = GpPkTiling(None, None, None, None, B, K, gpt, gpt * K)
    q_out = GpPkAdj(None, None, None, ap.data_ptr(), 0)
    q_in = GpPkAdj(None, None, None, ap.data_ptr(), 1)
    sq = _Stack(ws, B * K, wq, bq, _src(xp.data_ptr(), F, F), dbl, nd_e + nd_a, K)
    for l in range(Lq):
        _layer_fwd(tlq, q_out, None, K, [sq], l)
    zq = sq.concat()
    call('gp_pk_readout', C.byref(tlq), C.byref(zq), None, None, K, out.data_ptr(), arg.data_ptr(),
         C.c_longlong(ldo), F, st)
    ypred, acts = E.mlp_fwd(ws, out.data_ptr(), ldo, B, lin)
    tape = dict(B=B, N=N, K=K, rows=rows, F=F, Fa=Fa, ldo=ldo, dbl=dbl, gflat=gflat, goff=goff, shapes=shapes,
                imeta=imeta, subs=subs, cnt_pad=cnt_pad, info=info, ell=ell, ovf=ovf, rowmeta=rowmeta, xpack=xpack,
                axpack=axpack, tl1=tl1, tl2=tl2, tlq=tlq, a_out=a_out, a_in=a_in,
                q_out=q_out, q_in=q_in, se=se, sa=sa, sq=sq, pa=pa, S=S, xp=xp, ap=ap, out=out, arg=arg, acts=acts,
                lin=lin, x=x, adj=adj, assign_x=assign_x, nb=nb)
    return ypred, S, tape


def _gptr(tape, idx):
    return None if idx is None else tape['gflat'].data_ptr() + 4 * tape['goff'][idx]


def _layer_bwd(tape, tl, adj, adj_in, cnt_pad, Nn, specs):
    a = GpPkLayerBwdArgs()
    a.tl, a.adj, a.adj_in, a.cnt_pad, a.N, a.ns = tl, adj, adj_in, cnt_pad, Nn, len(specs)
    for i, sp in enumerate(specs):
        s, l = sp['stack'], sp['l']
        b = a.s[i]
        b.inp, b.out = s.src_in(l), s.src_out(l)
        b.rnorm, b.msums = s.rn[l].data_ptr(), s.msums[l]
        b.gl = sp['gl']
        b.W, b.b, b.dout = s.W[l].data_ptr(), _p(s.b[l]), s.d[l]
        iw, ib = sp['pair']
        b.dW, b.db = _gptr(tape, iw), _gptr(tape, ib)
        b.need_dx = int(sp['need_dx'])
        b.gz_prev = sp.get('gz_prev', GpPkGrad())
        b.gl_prev = sp.get('gl_prev')
        b.msums_prev = sp.get('msums_prev')
        b.dadj, b.dadj_acc = sp.get('dadj'), int(sp.get('dadj_acc', 0))
    call('gp_pk_layer_bwd', C.byref(a), _stream())


def backward(plan, tape, params, dypred, dS0):
    """Returns the parameter gradients as a list aligned with `params` (views of one flat buffer)."""
    st = _stream()
    ws = E.Workspace(tape['x'].device)
    B, N, K, F, Fa, ldo, rows = (tape[k] for k in ('B', 'N', 'K', 'F', 'Fa', 'ldo', 'rows'))
    se, sa, sq = tape['se'], tape['sa'], tape['sq']
    grads = [None] * len(params)
    if dypred is None:
        dypred = ws.z(B, plan.label_dim)
    dypred = E._chk(dypred, 'grad of ypred')
    dout = ws.f(B, ldo)
    gl = E.mlp_bwd(ws, dypred, B, tape['acts'], tape['lin'], dout.data_ptr(), ldo)
    for (iw, ib), (dw, db) in zip(plan.pred, gl):
        grads[iw] = dw
        if ib is not None:
            grads[ib] = db
    arg = tape['arg']

    # ---- pooled level
    dxp, dap = ws.f(B * K, F), ws.f(B, K, K)
    offq = sq.offs()
    glq = [ws.f(B * K, d) for d in sq.d[:-1]]
    for l in reversed(range(sq.L)):
        if l == sq.L - 1:
            g = GpPkGrad(None, 0, 0, dout.data_ptr(), arg.data_ptr(), ldo, F + offq[l])
        else:
            g = GpPkGrad(glq[l].data_ptr(), sq.d[l], 0, None, None, 0, 0)
        sp = dict(stack=sq, l=l, pair=plan.post[0][l], gl=g, need_dx=True, dadj=dap.data_ptr(),
                  dadj_acc=(l != sq.L - 1))
        if l > 0:
            sp['gz_prev'] = GpPkGrad(None, 0, 0, dout.data_ptr(), arg.data_ptr(), ldo, F + offq[l - 1])
            sp['gl_prev'] = glq[l - 1].data_ptr()
            sp['msums_prev'] = sq.msums[l - 1]
        else:
            sp['gl_prev'] = dxp.data_ptr()
        _layer_bwd(tape, tape['tlq'], tape['q_out'], tape['q_in'], None, K, [sp])

    # ---- pooling, assignment softmax / Linear
    gz, gza = ws.f(rows, F), ws.f(rows, Fa)
    pa = tape['pa']
    pa.dxp, pa.dap, pa.dout = dxp.data_ptr(), dap.data_ptr(), dout.data_ptr()
    pa.dS_ext = None if dS0 is None else E._chk(dS0, 'grad of assign_tensor').data_ptr()
    pa.gz, pa.gza = gz.data_ptr(), gza.data_ptr()
    iw, ib = plan.assign_pred[0]
    pa.dWp, pa.dbp = _gptr(tape, iw), _gptr(tape, ib)
    call('gp_pk_pool_bwd', C.byref(pa), st)

    # ---- level 0, both stacks in lock-step
    offe, offa = se.offs(), sa.offs()
    gle = [ws.f(rows, d) for d in se.d[:-1]]
    gla = [ws.f(rows, d) for d in sa.d[:-1]]
    cp = tape['cnt_pad'].data_ptr()
    for l in reversed(range(max(se.L, sa.L))):
        specs = []
        for s, pairs, gzb, Fw, off, glb in ((se, plan.emb, gz, F, offe, gle), (sa, plan.assign[0], gza, Fa, offa, gla)):
            if l >= s.L:
                continue
            if l == s.L - 1:
                g = GpPkGrad(gzb.data_ptr(), Fw, off[l], None, None, 0, 0)
            else:
                g = GpPkGrad(glb[l].data_ptr(), s.d[l], 0, None, None, 0, 0)
            sp = dict(stack=s, l=l, pair=pairs[l], gl=g, need_dx=l > 0)
            if l > 0:
                sp['gz_prev'] = GpPkGrad(gzb.data_ptr(), Fw, off[l - 1], None, None, 0, 0)
                sp['gl_prev'] = glb[l - 1].data_ptr()
                sp['msums_prev'] = s.msums[l - 1]
            specs.append(sp)
        _layer_bwd(tape, tape['tl1'], tape['a_out'], tape['a_in'], cp, N, specs)

    gflat, goff, shapes = tape['gflat'], tape['goff'], tape['shapes']
    done = set()
    for pairs in (plan.emb, plan.assign[0], plan.post[0], [plan.assign_pred[0]]):
        for iw, ib in pairs:
            for i in (iw, ib):
                if i is not None and i not in done:
                    grads[i] = gflat[goff[i]:goff[i + 1]].view(shapes[i])
                    done.add(i)
    return grads


# ------------------------------------------------------------------------------------------------------------------
# link-prediction loss on the packed level-0 blocks (encoders.py:1311-1331)
# ------------------------------------------------------------------------------------------------------------------
def link_forward(ws, S, adj, nb, inv, inv_dev, base):
    """total = base + link, link = inv * inv_dev * sum over the real blocks; returns (total, link) [1] tensors."""
    B, N, K = S.shape
    st = _stream()
    acc = ws.z(2)                                   # one double, zeroed
    call('gp_pk_link_fwd', S.data_ptr(), adj.data_ptr(), _p(nb), B, N, K, acc.data_ptr(), st)
    total, link = ws.f(1), ws.f(1)
    call('gp_pk_link_finalize', acc.data_ptr(), C.c_double(inv), _p(inv_dev), _p(base), total.data_ptr(),
         link.data_ptr(), st)
    return total, link


def link_backward(ws, S, adj, nb, alpha, alpha_dev, alpha_dev2):
    B, N, K = S.shape
    dS = ws.f(B, N, K)
    call('gp_pk_link_bwd', S.data_ptr(), adj.data_ptr(), _p(nb), B, N, K, C.c_float(alpha), _p(alpha_dev),
         _p(alpha_dev2), dS.data_ptr(), _stream())
    return dS
