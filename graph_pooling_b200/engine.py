"""Explicit forward/backward schedules of the DiffPool hot path over the C ABI (libgp_b200.so).

Nothing here computes: every tensor op is a call into the hand-written sm_100a kernels through
``_lib.call``.  PyTorch provides device buffers (``torch.empty``), the current stream and the
autograd graph edges only.  There is no CPU / eager fallback: a missing library or a non-CUDA
tensor raises.

Reference semantics followed (file:line under /root/reference):
  gcn stack      encoders.py:1054-1081   (GraphConv :315-328, ReLU+BN :1062-1064, concat :1078)
  readout        encoders.py:1097,1257,1287
  assignment     encoders.py:1269-1275
  pooling        encoders.py:1278-1279
  prediction     encoders.py:1021-1033,1299
  losses         encoders.py:1124-1127,1302-1334
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GpGemm, GpLayerBwd, call

F32 = 0  # gp_precision


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, name):
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: the gp_b200 hot path has no CPU fallback' % name)
    if t.dtype != torch.float32:
        raise ValueError('%s must be float32 (got %s)' % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _chk_adj(t, allow_u8):
    """adjacency: float32 (the reference feed, train.py:197) or -- tensor-core mode only -- uint8 {0,1} (compact
    feed: a quarter of the host->device bytes; expanded to the bf16 operand by gp_adj_prepare)."""
    if allow_u8 and t.dtype == torch.uint8:
        if not t.is_cuda:
            raise RuntimeError('adj must be a CUDA tensor: the gp_b200 hot path has no CPU fallback')
        return t if t.is_contiguous() else t.contiguous()
    return _chk(t, 'adj')


class Workspace:
    """Allocation helper: fresh torch buffers on the current device (caching allocator)."""

    def __init__(self, device):
        self.device = device

    def f(self, *shape):
        return torch.empty(shape, device=self.device, dtype=torch.float32)

    def z(self, *shape):
        t = torch.empty(shape, device=self.device, dtype=torch.float32)
        call('gp_fill_f32', t.data_ptr(), C.c_longlong(t.numel()), C.c_float(0.0), _stream())
        return t

    def i(self, *shape):
        return torch.empty(shape, device=self.device, dtype=torch.int32)


def bgemm(A, B, Cp, M, N, K, batch, sA, sB, sC, lim=None, lim_m=0, lim_n=0, lim_k=0, alpha=1.0, beta=0.0,
          alpha_dev=None, bias=None, relu=0, split_k=0):
    """C[b] = alpha*A[b].B[b] (+bias)(relu) + beta*C[b]; operands are raw device pointers,
    sA=(batch,m,k) sB=(batch,k,n) sC=(batch,m,n) element strides."""
    g = GpGemm(A, B, Cp, M, N, K, batch, sA[0], sA[1], sA[2], sB[0], sB[1], sB[2], sC[0], sC[1], sC[2],
               lim, lim_m, lim_n, lim_k, alpha, beta, alpha_dev, bias, relu, split_k)
    call('gp_bgemm_f32', C.byref(g), _stream())


# ------------------------------------------------------------------------------------------
# GCN stack  (encoders.py:1054-1081)
# ------------------------------------------------------------------------------------------
class StackCtx:
    __slots__ = ('B', 'N', 'din', 'douts', 'F', 'x_ptr', 'ldx', 'adj', 'nb', 'weights', 'biases', 'add_self',
                 'bn', 'zcat', 'layers', 'keep', 'drops')


def layer_seed(seed, l):
    """Per-layer dropout seed derived from the forward call's base seed (also used by the tests to rebuild a mask)."""
    return (int(seed) + 0x632BE5AB * (l + 1)) & ((1 << 62) - 1)


def dropout(x_ptr, ldx, rows, d, p, seed, y_ptr, ldy, yb_ptr=None, ldyb=0):
    call('gp_dropout_f32', x_ptr, C.c_longlong(ldx), C.c_longlong(rows), d, C.c_float(p), C.c_ulonglong(seed),
         y_ptr, C.c_longlong(ldy), yb_ptr, C.c_longlong(ldyb), _stream())


def stack_forward(ws, x_ptr, ldx, din, adj, nb, B, N, weights, biases, add_self, bn, prec, keep=(), drops=None,
                  seed=0):
    """Runs L GraphConv layers (+ReLU+BN between) and returns the UNMASKED concat buffer
    zcat [B,N,F]; the mask of encoders.py:1078-1080 is applied by the consumers (readout treats
    pad rows as 0; pooling multiplies by S whose pad rows are 0)."""
    st = _stream()
    L = len(weights)
    douts = [int(w.shape[1]) for w in weights]
    Fw = sum(douts)
    zcat = ws.f(B, N, Fw)
    ctx = StackCtx()
    ctx.B, ctx.N, ctx.din, ctx.douts, ctx.F = B, N, din, douts, Fw
    ctx.x_ptr, ctx.ldx, ctx.adj, ctx.nb = x_ptr, ldx, adj, nb
    ctx.weights, ctx.biases, ctx.add_self, ctx.bn = weights, biases, add_self, bn
    ctx.zcat, ctx.layers, ctx.keep = zcat, [], list(keep)
    ctx.drops = [None] * L
    cur_ptr, cur_ld, cur_d, off = x_ptr, ldx, din, 0
    zp = zcat.data_ptr()
    for l in range(L):
        last = l == L - 1
        dout = douts[l]
        if drops is not None and drops[l] > 0.0:          # nn.Dropout on this layer's input (encoders.py:316-317)
            xd = ws.f(B, N, cur_d)
            ctx.drops[l] = (float(drops[l]), layer_seed(seed, l))
            dropout(cur_ptr, cur_ld, B * N, cur_d, ctx.drops[l][0], ctx.drops[l][1], xd.data_ptr(), cur_d)
            ctx.keep.append(xd)
            cur_ptr, cur_ld = xd.data_ptr(), cur_d
        u = ws.f(B, N, cur_d)
        rnorm = ws.f(B, N)
        slot = zp + off * 4
        if last:
            y, y_ptr, ldy = None, slot, Fw
        else:
            y = ws.f(B, N, dout)
            y_ptr, ldy = y.data_ptr(), dout
        call('gp_graphconv_fwd', cur_ptr, cur_ld, _p(adj), _p(weights[l]), _p(biases[l]), _p(nb), B, N, cur_d, dout,
             int(add_self), 1, _p(u), y_ptr, ldy, _p(rnorm), prec, st)
        mean = invstd = None
        if not last:
            if bn:
                mean, invstd = ws.f(N), ws.f(N)
            nws = int(_lib.load().gp_relu_bn_fwd_ws(B, N, dout)) if bn else 0
            call('gp_relu_bn_fwd_x', y_ptr, slot, Fw, _p(mean), _p(invstd), B, N, dout, 1, int(bn),
                 _p(ws.f(nws)) if nws else None, st)
        ctx.layers.append((cur_ptr, cur_ld, cur_d, dout, off, u, y, rnorm, mean, invstd))
        cur_ptr, cur_ld, cur_d = slot, Fw, dout
        off += dout
    return zcat, ctx


def stack_backward(ws, ctx, dz_ptr, lddz, dout_ptr, arg_ptr, ldo, need_dx, dadj, prec):
    """Backward of stack_forward.  dz: dense gradient of zcat (pointer, row stride) or None;
    dout/arg: max-readout gradient and winners for this stack's F columns or None.
    Returns ([(dW, db)] per layer, dx or None); dA is accumulated into `dadj` if given."""
    st = _stream()
    B, N, Fw = ctx.B, ctx.N, ctx.F
    L = len(ctx.layers)
    grads = [None] * L
    dxn = None
    zp = ctx.zcat.data_ptr()
    for l in reversed(range(L)):
        x_ptr, ldx, din, dout, off, u, y, rnorm, mean, invstd = ctx.layers[l]
        last = l == L - 1
        slot = zp + off * 4
        dv = ws.f(B, N, dout)
        q = GpLayerBwd()
        q.dz, q.lddz = (None if dz_ptr is None else dz_ptr + off * 4), lddz
        q.dxn, q.lddxn = _p(dxn), 0
        q.dout = None if dout_ptr is None else dout_ptr + off * 4
        q.argidx = None if arg_ptr is None else arg_ptr + off * 4
        q.ldo = ldo
        q.h, q.ldh = slot, Fw
        q.y, q.ldy = (slot, Fw) if last else (_p(y), dout)
        q.rnorm, q.mean, q.invstd = _p(rnorm), None, _p(invstd)
        q.B, q.N, q.d = B, N, dout
        q.relu, q.bn, q.normalize = int(not last), int(ctx.bn and not last), 1
        q.dv, q.dv_bf16, q.lddvb, q.db, q.ws = _p(dv), None, 0, None, None
        wsf = ws.f(int(_lib.load().gp_gcn_layer_bwd_ws_x(C.byref(q))))
        q.ws = _p(wsf)
        call('gp_gcn_layer_bwd_x', C.byref(q), st)
        need_dx_l = need_dx or l > 0
        w = ctx.weights[l]
        if ctx.biases[l] is not None:       # dW | db contiguous: the per-graph fused path reduces both in one pass
            dwdb = ws.f(din * dout + dout)
            dw, db = dwdb[:din * dout].view(din, dout), dwdb[din * dout:]
        else:
            dw, db = ws.f(din, dout), None
        du = ws.f(B, N, din) if (need_dx_l or dadj is not None) else None
        dx = ws.f(B, N, din) if need_dx_l else None
        cs = ws.f(int(_lib.load().gp_graphconv_bwd_ws(B, N, din, dout, int(ctx.add_self))))
        call('gp_graphconv_bwd', _p(dv), _p(u), x_ptr, ldx, _p(ctx.adj), _p(w), _p(ctx.nb), B, N, din, dout,
             int(ctx.add_self), _p(dw), _p(db), _p(du), _p(dx), _p(dadj), _p(cs), prec, st)
        grads[l] = (dw, db)
        if dx is not None and ctx.drops[l] is not None:   # gradient of the dropped input -> gradient of the input
            pd, sd = ctx.drops[l]
            dropout(dx.data_ptr(), din, B * N, din, pd, sd, dx.data_ptr(), din)
        dxn = dx
    return grads, dxn


# ------------------------------------------------------------------------------------------
# nn.Linear chains (pred_model, assign_pred) on the library's own GEMM
# ------------------------------------------------------------------------------------------
def linear_fwd(ws, x_ptr, ldx, rows, w, b, relu):
    out_f, in_f = int(w.shape[0]), int(w.shape[1])
    y = ws.f(rows, out_f)
    bgemm(x_ptr, _p(w), _p(y), rows, out_f, in_f, 1, (0, ldx, 1), (0, 1, in_f), (0, out_f, 1),
          bias=_p(b), relu=int(relu))
    return y


def linear_bwd(ws, dy, x_ptr, ldx, rows, w, has_bias, need_dx, dx_ptr=None, lddx=None):
    """Returns (dW, db, dx); with dx_ptr the input gradient is written there (row stride lddx)."""
    out_f, in_f = int(w.shape[0]), int(w.shape[1])
    dw = ws.f(out_f, in_f)
    split = max(1, min(512, rows // 1024))     # split-K over the rows: enough CTAs to stream them at HBM rate
    bgemm(_p(dy), x_ptr, _p(dw), out_f, in_f, rows, 1, (0, 1, out_f), (0, ldx, 1), (0, in_f, 1), split_k=split)
    db = None
    if has_bias:
        db = ws.f(out_f)
        cs = ws.f(256 * out_f)
        call('gp_colsum_f32', _p(dy), C.c_longlong(rows), out_f, C.c_longlong(out_f), _p(db), 0, _p(cs), _stream())
    dx = None
    if need_dx:
        if dx_ptr is None:
            dx = ws.f(rows, in_f)
            dx_ptr, lddx = dx.data_ptr(), in_f
        bgemm(_p(dy), _p(w), dx_ptr, rows, in_f, out_f, 1, (0, out_f, 1), (0, in_f, 1), (0, lddx, 1))
    return dw, db, dx


def mlp_fwd(ws, x_ptr, ldx, rows, linears):
    """pred_model: Linear (+ReLU between), encoders.py:1021-1033.  Returns (ypred, saved acts)."""
    acts = [(x_ptr, ldx, None)]
    ptr, ld, h = x_ptr, ldx, None
    for i, (w, b) in enumerate(linears):
        h = linear_fwd(ws, ptr, ld, rows, w, b, relu=(i < len(linears) - 1))
        ptr, ld = h.data_ptr(), h.shape[1]
        # the last activation is the encoder's OUTPUT (ypred): the tape keeps a detached alias, never the output
        # object itself (that would close a reference cycle through the autograd node)
        acts.append((ptr, ld, h.detach()))
    return h, acts


def mlp_bwd(ws, dy, rows, acts, linears, dx_ptr, lddx):
    """Backward of mlp_fwd; the gradient of the MLP input is written to (dx_ptr, lddx)."""
    grads = [None] * len(linears)
    g = dy
    for i in reversed(range(len(linears))):
        w, b = linears[i]
        if i < len(linears) - 1:          # ReLU after layer i
            a = acts[i + 1][2]
            gm = ws.f(*a.shape)
            call('gp_relu_mask_bwd', _p(g), _p(a), C.c_longlong(g.numel()), _p(gm), _stream())
            g = gm
        xp_, xld, _ = acts[i]
        if i == 0:
            dw, db, dx = linear_bwd(ws, g, xp_, xld, rows, w, b is not None, True, dx_ptr, lddx)
        else:
            dw, db, dx = linear_bwd(ws, g, xp_, xld, rows, w, b is not None, True)
        grads[i] = (dw, db)
        g = dx
    return grads


def prep_nb(batch_num_nodes, N, device):
    """batch_num_nodes (host numpy / list / tensor, train.py:200) -> (int32 device tensor, host array)."""
    if batch_num_nodes is None:
        return None, None
    if torch.is_tensor(batch_num_nodes) and batch_num_nodes.is_cuda:
        # device-resident node counts (int32): used as they are, no host copy, no synchronisation -- what a captured
        # CUDA graph needs (graphed.py).  The host never sees the values: range checks are the caller's job.
        if batch_num_nodes.dtype != torch.int32 or batch_num_nodes.dim() != 1:
            raise ValueError('device batch_num_nodes must be a 1-D int32 tensor')
        return batch_num_nodes.contiguous(), None
    if torch.is_tensor(batch_num_nodes):
        host = batch_num_nodes.detach().cpu().numpy()
    else:
        host = np.asarray(batch_num_nodes)
    host = np.ascontiguousarray(host.astype(np.int32))
    if host.ndim != 1:
        raise ValueError('batch_num_nodes must be 1-D')
    if host.size and (host.min() < 0 or host.max() > N):
        raise ValueError('batch_num_nodes out of range [0, %d]' % N)
    # one small pinned transfer: [n_b | argsort(-n_b)] -- the second half lets the persistent GEMMs walk ragged
    # batches from the largest graph to the smallest (balanced static schedule)
    both = np.concatenate([host, np.argsort(-host.astype(np.int64), kind='stable').astype(np.int32)])
    both_dev = torch.from_numpy(both).pin_memory().to(device, non_blocking=True)
    dev, order = both_dev[:host.size], both_dev[host.size:]
    if host.size and int(host.min()) != int(host.max()):
        from . import engine_tc
        engine_tc.register_order(dev, order)
        dev._gp_order = order                       # keep the permutation alive as long as the node counts
    return dev, host
