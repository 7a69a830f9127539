"""Tensor-core (GP_BF16) schedule of the DiffPool path: every dense contraction runs on
gp_bgemm_bf16 (TMA + tcgen05.mma, fp32 accumulation in TMEM); row/node reductions stay fp32.

Operands are bf16 shadow buffers with row strides padded to 8 elements (TMA's 16-byte rule);
they are produced either by a GEMM epilogue (bf16 copy output) or by gp_cvt_f32_bf16.  All
products read their operands in natural row-major layout (K-major / MN-major UMMA descriptors),
so nothing is ever transposed in HBM.  Same reference semantics as engine.py.
"""
import ctypes as C
import os

import torch

from . import engine as E
from ._lib import GpGemmBf16x, GpLayerBwd, call, load

BF16 = 1
KM, MN = 0, 1        # operand major-ness (see include/gp_b200.h)


def r8(n):
    return (int(n) + 7) & ~7


class Op:
    """bf16 operand view: device pointer, row stride, batch stride (elements), keep-alive tensor."""
    __slots__ = ('ptr', 'ld', 'sb', 't')

    def __init__(self, ptr, ld, sb, t=None):
        self.ptr, self.ld, self.sb, self.t = ptr, ld, sb, t


def bfbuf(ws, B, R, Ccols, pad=8):
    """bf16 operand buffer [B, R, ld], ld = Ccols rounded up to `pad` elements (8: TMA's 16-byte rule; 32: adjacency-sized
    operands whose rows the link-loss epilogue reads / writes in whole 32-column chunks with 256-bit accesses)."""
    ld = (int(Ccols) + pad - 1) // pad * pad
    t = torch.empty(B, R, ld, device=ws.device, dtype=torch.bfloat16)
    return Op(t.data_ptr(), ld, R * ld, t)


def cvt(ws, x_ptr, ldx, rows, cols, B=1, out=None):
    """fp32 [rows, cols] (row stride ldx) -> bf16 operand [B, rows/B, r8(cols)]."""
    if out is None:
        out = bfbuf(ws, B, rows // B, cols)
    call('gp_cvt_f32_bf16', x_ptr, ldx, out.ptr, out.ld, C.c_longlong(rows), cols, min(r8(cols), out.ld),
         E._stream())
    return out


def adj_prepare(ws, adj, nb, B, N):
    """adj [B,N,N] fp32 (or uint8) -> (bf16 operand, flags int32[2] on device: [not symmetric, not {0,1}])."""
    out = bfbuf(ws, B, N, N, pad=32)
    flags = torch.empty(2, device=ws.device, dtype=torch.int32)
    call('gp_adj_prepare', adj.data_ptr(), 1 if adj.dtype == torch.uint8 else 0, E._p(nb), B, N, out.ptr, out.ld,
         flags.data_ptr(), E._stream())
    return out, flags


# device pointer of a node-count vector -> (argsort(-n_b) on the device, batch size), registered by engine.prep_nb for
# host-side node counts.  The entry holds the permutation tensor -- which shares its storage with the node counts --
# so the key address cannot be reused while the entry exists.  Batched contractions clipped by that vector walk the
# graphs longest-first; any permutation is a valid schedule, so a stale entry of the same length is harmless.
_ORDER = {}


def register_order(nb_dev, order_dev):
    if len(_ORDER) > 64:
        _ORDER.clear()
    _ORDER[nb_dev.data_ptr()] = (order_dev, int(order_dev.numel()))


class PreparedAdjacency:
    """The bf16 adjacency operand of a batch, built by the FEED instead of by the module: accepted wherever the
    encoders (tensor-core mode) take `adj`.  A batch may be assembled from parts that reached the device in different
    encodings -- fp32 (the reference feed, train.py:197), uint8, or bit-packed rows (gp_host_pack_adj_bits) -- each
    expanded by gp_adj_prepare_x into its slice of the operand; the symmetry / {0,1} flags accumulate over the parts."""
    KINDS = {'f32': 0, 'u8': 1, 'bits': 2}

    def __init__(self, B, N, device):
        self.shape = (B, N, N)
        self.device = torch.device(device)
        self.op = bfbuf(E.Workspace(self.device), B, N, N, pad=32)
        self.flags = torch.zeros(2, device=self.device, dtype=torch.int32)
        self._fresh = True

    def reset(self):
        self._fresh = True

    def add(self, src, b0, nb=None, kind='f32'):
        """Graphs [b0, b0 + src.shape[0]) from `src`: [cnt,N,N] fp32 / uint8, or [cnt,N,ldb] uint8 bit rows."""
        B, N, _ = self.shape
        cnt = int(src.shape[0])
        if not src.is_cuda or not src.is_contiguous() or b0 < 0 or b0 + cnt > B:
            raise ValueError('PreparedAdjacency.add: need a contiguous CUDA tensor inside the batch')
        code = self.KINDS[kind]
        ld_in = int(src.shape[2])
        nbp = None if nb is None else nb.data_ptr() + 4 * b0
        call('gp_adj_prepare_x', src.data_ptr(), code, C.c_longlong(ld_in), nbp, cnt, N,
             self.op.ptr + 2 * b0 * self.op.sb, self.op.ld, self.flags.data_ptr(), 0 if self._fresh else 1, E._stream())
        self._fresh = False
        return self

    def from_edges(self, edges, eptr, max_edges_per_graph, undirected=True):
        """The whole batch from its edge lists (device tensors: edges [E,2] graph-local ids, int32 or 16-bit; eptr [B+1] int32): the
        adjacency never exists as fp32 / uint8 anywhere (gp_adj_from_edges)."""
        B, N, _ = self.shape
        if edges.dtype not in (torch.int32, torch.int16, torch.uint16) or eptr.dtype != torch.int32 or \
                not edges.is_cuda or not eptr.is_cuda:
            raise ValueError('PreparedAdjacency.from_edges: need CUDA tensors, edges int32 or 16-bit, eptr int32')
        if eptr.numel() != B + 1 or not edges.is_contiguous():
            raise ValueError('PreparedAdjacency.from_edges: eptr must have B + 1 entries, edges must be contiguous')
        call('gp_adj_from_edges', edges.data_ptr(), int(edges.element_size()), eptr.data_ptr(), B, N,
             int(max_edges_per_graph), int(undirected), self.op.ptr, self.op.ld, self.flags.data_ptr(), 0, E._stream())
        self._fresh = False
        return self


def tcgemm_multi(pairs, M, N, batch, Cf=None, Cb=None, lim=None, lim_m=0, lim_n=0, alpha=1.0, beta=0.0,
                 alpha_dev=None, bias=None, relu=0, split_k=0, cond=None, cond_npairs=0, cond_alpha=1.0, tri=0):
    """One persistent tcgen05 launch accumulating sum_q A_q.B_q (gp_bgemm_bf16x).
    pairs: [(A Op, a_major, B Op, b_major, K, lim_k)]; Cf = (ptr, ld, sb) fp32 out, Cb = Op bf16 out."""
    g = GpGemmBf16x()
    g.npairs = len(pairs)
    for q, (A, am, Bo, bm, K, lk) in enumerate(pairs):
        pr = g.pair[q]
        pr.A, pr.B, pr.K = A.ptr, Bo.ptr, K
        pr.ldA, pr.sAb, pr.a_major = A.ld, A.sb, am
        pr.ldB, pr.sBb, pr.b_major = Bo.ld, Bo.sb, bm
        pr.lim_k = lk
    cp, cld, csb = Cf if Cf is not None else (None, 0, 0)
    g.C, g.Cb = cp, (None if Cb is None else Cb.ptr)
    g.M, g.N, g.batch = M, N, batch
    g.ldC, g.sCb = cld, csb
    g.ldCb, g.sCbb = (0, 0) if Cb is None else (Cb.ld, Cb.sb)
    g.lim, g.lim_m, g.lim_n = lim, lim_m, lim_n
    g.alpha, g.beta, g.alpha_dev = alpha, beta, alpha_dev
    g.bias, g.relu, g.split_k = bias, relu, split_k
    g.cond, g.cond_npairs, g.cond_alpha = cond, cond_npairs, cond_alpha
    g.tri = tri
    ent = _ORDER.get(lim) if lim is not None else None
    g.order = ent[0].data_ptr() if (ent is not None and ent[1] == batch) else None   # longest-first walk (ragged)
    call('gp_bgemm_bf16x', C.byref(g), E._stream())


def tcgemm(A, a_major, Bo, b_major, M, N, K, batch, Cf=None, Cb=None, lim=None, lim_m=0, lim_n=0, lim_k=0,
           alpha=1.0, beta=0.0, alpha_dev=None, bias=None, relu=0, split_k=0, cond=None, cond_npairs=0):
    """Single product on tensor cores.  Cf = (ptr, ld, sb) fp32 output, Cb = Op bf16 output."""
    return tcgemm_multi([(A, a_major, Bo, b_major, K, lim_k)], M, N, batch, Cf, Cb, lim, lim_m, lim_n, alpha,
                        beta, alpha_dev, bias, relu, split_k, cond, cond_npairs)


def pick_split(M, N, K):
    """Split-K factor of a weight-gradient GEMM (few output tiles, a contraction over all B*N rows): the one whose work
    items fill the persistent grid best in at most three rounds.  The grid is 148 CTAs, or 74 CTA pairs when the launch is
    pair-eligible (gemm_tc2.cu `run`: 256-column tiles, >= 2 row blocks, last 256-row block not mostly padding) -- e.g. cfg4's
    dWp (512 x 768): 6 pair tiles, 37 splits = 222 items = exactly three rounds (25 splits were 2.03 rounds: 68 %)."""
    t128, t256 = (M + 127) // 128, (M + 255) // 256
    pair = N > 128 and t128 >= 2 and 2 * t256 * 16 <= t128 * 17 and not os.environ.get('GP_NO_PAIR')
    tiles = (t256 if pair else t128) * ((N + 255) // 256 if N > 128 else 1)
    units = 74 if pair else 148
    smax = int(max(1, min(K // 256, 1024)))              # at least four 64-wide k-steps per split
    best, best_eff = 1, 0.0
    for sp in range(1, min(smax, (3 * units) // tiles + 1) + 1):
        items = tiles * sp
        rounds = -(-items // units)
        if rounds > 3:
            break
        eff = items / float(units * rounds)
        if eff > best_eff + 1e-9:
            best, best_eff = sp, eff
    return best


# ------------------------------------------------------------------------------------------
class StackCtxTC:
    pass


def _norm_gemm(A, Bo, M, N, K, bias, Cf, Cb, rnorm, rowstat, stat_relu):
    """V = A.B + bias, Y = V / max(||V||, eps) in the GEMM epilogue (gp_bgemm_bf16_norm); A K-major, B N-major."""
    g = GpGemmBf16x()
    g.npairs = 1
    pr = g.pair[0]
    pr.A, pr.B, pr.K = A.ptr, Bo.ptr, K
    pr.ldA, pr.sAb, pr.a_major = A.ld, 0, KM
    pr.ldB, pr.sBb, pr.b_major = Bo.ld, 0, MN
    pr.lim_k = 0
    cp, cld, _ = Cf if Cf is not None else (None, 0, 0)
    g.C, g.Cb = cp, (None if Cb is None else Cb.ptr)
    g.M, g.N, g.batch = M, N, 1
    g.ldC, g.sCb = cld, 0
    g.ldCb, g.sCbb = (0, 0) if Cb is None else (Cb.ld, 0)
    g.lim, g.lim_m, g.lim_n = None, 0, 0
    g.alpha, g.beta, g.alpha_dev = 1.0, 0.0, None
    g.bias, g.relu, g.split_k = bias, 0, 0
    g.cond, g.cond_npairs, g.cond_alpha = None, 0, 1.0
    g.order = None
    call('gp_bgemm_bf16_norm', C.byref(g), rnorm, rowstat, int(stat_relu), E._stream())


def _stack_begin(ws, xb, din, adjb, nb, B, N, weights, biases, bn, pad_last=0):
    """pad_last: run the LAST layer at this width (> its weight's column count): the extra output columns have zero
    weights and zero bias, so they are exact zeros everywhere (V = 0, Y = 0/||V|| = 0, dV = 0) and every buffer of the
    stack keeps 16-byte-aligned rows.  Used for assignment GCNs whose cluster count K is not a multiple of 8."""
    L = len(weights)
    wcols = [int(w.shape[1]) for w in weights]
    douts = list(wcols)
    if pad_last:
        douts[-1] = int(pad_last)
    Fw = sum(douts)
    ctx = StackCtxTC()
    ctx.B, ctx.N, ctx.douts, ctx.F, ctx.adjb, ctx.nb, ctx.bn = B, N, douts, Fw, adjb, nb, bn
    ctx.wcols = wcols
    ctx.weights, ctx.biases, ctx.layers = weights, biases, []
    ctx.zcat = ws.f(B, N, Fw)
    ctx.zb = bfbuf(ws, B, N, Fw)
    ctx.offs = [sum(douts[:l]) for l in range(L)]
    ctx.aligned = all(o % 8 == 0 for o in ctx.offs)
    ctx.h32 = True           # write H = BN(relu(Y)) of the non-last layers as fp32 too (readout / dropout / unaligned concat)
    ctx.cur, ctx.cur_d = xb, din
    return ctx


def _layer_forward(ws, ctx, l, ub, hb2=None):
    """Layer l of a stack given its U = A.X operand `ub` (bf16 Op, possibly a column half of a wider buffer):
    Y = normalize(U.W + b) with the row norm and the BatchNorm row sums taken in the GEMM epilogue, then
    H = BN(relu(Y)) written once as fp32 (concat slot) and bf16 (concat operand slot, plus `hb2` if given)."""
    st = E._stream()
    B, N, Fw, zb = ctx.B, ctx.N, ctx.F, ctx.zb
    L = len(ctx.weights)
    last = l == L - 1
    dout, off = ctx.douts[l], ctx.offs[l]
    cur, cur_d = ctx.cur, ctx.cur_d
    w = ctx.weights[l]
    rows = B * N
    zp = ctx.zcat.data_ptr()
    wc = ctx.wcols[l]
    wb = cvt(ws, w.data_ptr(), wc, cur_d, wc)                                       # [din, r8(wc)], zero pad columns
    bias_p = E._p(ctx.biases[l])
    if wc != dout and bias_p is not None:                                           # padded last layer: zero bias beyond wc
        bpad = ws.f(dout)
        call('gp_pad_copy_f32', bias_p, C.c_longlong(wc), C.c_longlong(1), wc, bpad.data_ptr(), C.c_longlong(dout),
             C.c_longlong(1), dout, C.c_float(0.0), st)
        bias_p = bpad.data_ptr()
    slot = zp + off * 4
    # bf16 operand copy of this layer's output: a column slot of zb when 16-byte aligned, else its own buffer
    hb = Op(zb.ptr + off * 2, zb.ld, zb.sb, zb.t) if ctx.aligned else bfbuf(ws, B, N, dout)
    hb_flat = Op(hb.ptr, hb.ld, 0)
    if last:
        y, y_ptr, ldy = None, slot, Fw
    else:
        y = ws.f(B, N, dout)
        y_ptr, ldy = y.data_ptr(), dout
    rnorm = ws.f(B, N)
    use_bn = bool(ctx.bn and not last)
    uflat, wflat = Op(ub.ptr, ub.ld, 0), Op(wb.ptr, wb.ld, 0)
    mean = invstd = None
    if dout <= 512:                                      # rows up to 512 wide stay in TMEM for the fused tail
        rowstat = ws.f(rows, 2) if use_bn else None
        _norm_gemm(uflat, wflat, rows, dout, cur_d, bias_p, (y_ptr, ldy, 0), hb_flat if last else None,
                   rnorm.data_ptr(), E._p(rowstat), 1)
        if use_bn:
            mean, invstd = ws.f(N), ws.f(N)
            call('gp_bn_finalize', rowstat.data_ptr(), B, N, dout, mean.data_ptr(), invstd.data_ptr(), st)
        if not last:
            call('gp_bn_apply', y_ptr, ldy, E._p(mean), E._p(invstd), B, N, dout, 1, int(use_bn),
                 slot if (ctx.h32 or not ctx.aligned) else None, Fw,
                 hb.ptr, hb.ld, None if hb2 is None else hb2.ptr, 0 if hb2 is None else hb2.ld, st)
    else:
        # wide layer (e.g. the assignment GCN's last layer, dout = K): plain GEMM, then one normalize pass
        tcgemm(uflat, KM, wflat, MN, rows, dout, cur_d, 1, Cf=(y_ptr, ldy, 0), bias=bias_p)
        yb_ok = dout % 4 == 0 and dout <= 2048 and ldy % 4 == 0
        call('gp_bias_normalize_x', y_ptr, None, rnorm.data_ptr(), C.c_longlong(rows), dout, ldy, 1,
             hb.ptr if (last and yb_ok) else None, hb.ld, st)
        if last and not yb_ok:
            cvt(ws, slot, Fw, rows, dout, out=hb_flat)
        if not last:
            if use_bn:
                mean, invstd = ws.f(N), ws.f(N)
            call('gp_relu_bn_fwd', y_ptr, slot, Fw, E._p(mean), E._p(invstd), B, N, dout, 1, int(use_bn), st)
            cvt(ws, slot, Fw, rows, dout, out=hb_flat)
            if hb2 is not None:
                cvt(ws, slot, Fw, rows, dout, out=Op(hb2.ptr, hb2.ld, 0))
    ctx.layers.append((cur, cur_d, dout, off, ub, y, rnorm, mean, invstd, wb))
    ctx.cur, ctx.cur_d = hb, dout


def _stack_end(ws, ctx):
    if not ctx.aligned:
        cvt(ws, ctx.zcat.data_ptr(), ctx.F, ctx.B * ctx.N, ctx.F, out=ctx.zb)
    return ctx.zcat, ctx.zb, ctx


def stack_forward(ws, xb, din, adjb, nb, B, N, weights, biases, bn, u0=None, pad_last=0, drops=None, seed=0, h32=True):
    """TC version of engine.stack_forward (add_self unsupported).  xb/adjb: bf16 operands.
    Per layer:  U = A.X (tcgen05)  ->  _layer_forward.
    u0: optional precomputed U of the first layer (shared with another stack that has the same A and X).
    Returns (zcat fp32 [B,N,F], zb bf16 operand of the same concat, ctx)."""
    ctx = _stack_begin(ws, xb, din, adjb, nb, B, N, weights, biases, bn, pad_last)
    # h32 = False (assignment GCNs: nobody reads the fp32 H of their inner layers -- the assignment head, the next layer
    # and the backward all use the bf16 operand / the saved Y): 4 bytes per element less to write per inner layer
    ctx.h32 = bool(h32) or (drops is not None and any(d > 0.0 for d in drops))
    nbp, lim = E._p(nb), int(nb is not None)
    ctx.drops = [None] * len(weights)
    for l in range(len(weights)):
        if drops is not None and drops[l] > 0.0 and l > 0:
            # nn.Dropout on this layer's input (encoders.py:316-317): the dropped bf16 operand is made from the fp32
            # concat slot of the previous layer; the backward re-applies the mask to dX from the same seed
            ctx.drops[l] = (float(drops[l]), E.layer_seed(seed, l))
            xdb = bfbuf(ws, B, N, ctx.cur_d)
            E.dropout(ctx.zcat.data_ptr() + ctx.offs[l - 1] * 4, ctx.F, B * N, ctx.cur_d, ctx.drops[l][0],
                      ctx.drops[l][1], None, 0, xdb.ptr, xdb.ld)
            ctx.cur = xdb
        if l == 0 and u0 is not None:
            ub = u0
        else:
            ub = bfbuf(ws, B, N, ctx.cur_d)
            # U = A.X : A K-major, X N-major
            tcgemm(adjb, KM, ctx.cur, MN, N, ctx.cur_d, N, B, Cb=ub, lim=nbp, lim_m=lim, lim_k=lim)
        _layer_forward(ws, ctx, l, ub)
    return _stack_end(ws, ctx)


def dual_ok(wE, wA):
    """Two GCN stacks over the same adjacency can run in lock-step (one A.X per layer for both) when they have the
    same depth and their hidden widths keep both column halves 16-byte aligned and inside the fused-epilogue limit."""
    if len(wE) != len(wA) or len(wE) < 2:
        return False
    for l in range(len(wE) - 1):
        de, da = int(wE[l].shape[1]), int(wA[l].shape[1])
        if de % 8 or da % 8 or de > 256 or da > 256:
            return False
    return True


def dual_stack_forward(ws, xb, din, xab, dina, adjb, nb, B, N, wE, bE, bnE, wA, bA, bnA, pad_lastA=0):
    """Embedding GCN and assignment GCN of one level in lock-step: both multiply the SAME
    adjacency, so layer l's two inputs sit side by side in one [B,N,He+Ha] operand and ONE pass over A produces
    both U's (A.X at 128 columns is HBM-bound on reading A: sharing the pass halves that traffic)."""
    cE = _stack_begin(ws, xb, din, adjb, nb, B, N, wE, bE, bnE)
    cA = _stack_begin(ws, xab, dina, adjb, nb, B, N, wA, bA, bnA, pad_lastA)
    cA.h32 = False           # the assignment GCN's inner H is consumed as bf16 only
    nbp, lim = E._p(nb), int(nb is not None)
    L = len(wE)
    hcat = None
    for l in range(L):
        if l == 0:
            uE = bfbuf(ws, B, N, din)
            tcgemm(adjb, KM, xb, MN, N, din, N, B, Cb=uE, lim=nbp, lim_m=lim, lim_k=lim)
            if xab is xb:                                # same A, same X: one U for both first layers
                uA = uE
            else:
                uA = bfbuf(ws, B, N, dina)
                tcgemm(adjb, KM, xab, MN, N, dina, N, B, Cb=uA, lim=nbp, lim_m=lim, lim_k=lim)
        else:
            de, da = cE.douts[l - 1], cA.douts[l - 1]
            ucat = bfbuf(ws, B, N, de + da)
            tcgemm(adjb, KM, hcat, MN, N, de + da, N, B, Cb=ucat, lim=nbp, lim_m=lim, lim_k=lim)
            uE = Op(ucat.ptr, ucat.ld, ucat.sb, ucat.t)
            uA = Op(ucat.ptr + de * 2, ucat.ld, ucat.sb, ucat.t)
        h2E = h2A = None
        if l < L - 1:
            de, da = cE.douts[l], cA.douts[l]
            hcat = bfbuf(ws, B, N, de + da)              # next layer's side-by-side input, filled by both tails
            h2E = Op(hcat.ptr, hcat.ld, hcat.sb, hcat.t)
            h2A = Op(hcat.ptr + de * 2, hcat.ld, hcat.sb, hcat.t)
        _layer_forward(ws, cE, l, uE, h2E)
        _layer_forward(ws, cA, l, uA, h2A)
    return _stack_end(ws, cE), _stack_end(ws, cA)


def _layer_backward_head(ws, ctx, l, dz_ptr, lddz, dxn, lddxn, dout_ptr, arg_ptr, ldo, dz_bf16=False, dxn_bf16=False):
    """Element-wise tail backward of layer l (bf16 dV + db in ONE pass over HBM; Hhat is recomputed from Y and the
    saved statistics) and dW = U^T dV.  Returns (dvb, dw, db)."""
    st = E._stream()
    B, N, Fw = ctx.B, ctx.N, ctx.F
    L = len(ctx.layers)
    xb, din, dout, off, ub, y, rnorm, mean, invstd, wb = ctx.layers[l]
    last = l == L - 1
    rows = B * N
    slot = ctx.zcat.data_ptr() + off * 4
    dvb = bfbuf(ws, 1, rows, dout)
    wc = ctx.wcols[l]                                    # < dout for a padded last layer (its dV pad columns are 0)
    has_b = ctx.biases[l] is not None
    db = ws.f(dout) if has_b else None
    q = GpLayerBwd()
    q.dz, q.lddz = (None if dz_ptr is None else dz_ptr + off * (2 if dz_bf16 else 4)), lddz
    q.dxn, q.lddxn = dxn, lddxn
    q.dz_bf16, q.dxn_bf16 = int(bool(dz_bf16 and dz_ptr is not None)), int(bool(dxn_bf16 and dxn is not None))
    q.dout = None if dout_ptr is None else dout_ptr + off * 4
    q.argidx = None if arg_ptr is None else arg_ptr + off * 4
    q.ldo = ldo
    use_bn = bool(ctx.bn and not last)
    q.h, q.ldh = None, Fw
    q.y, q.ldy = (slot, Fw) if last else (E._p(y), dout)
    q.rnorm, q.mean, q.invstd = E._p(rnorm), E._p(mean), E._p(invstd)
    q.B, q.N, q.d = B, N, dout
    q.relu, q.bn, q.normalize = int(not last), int(use_bn), 1
    q.dv, q.dv_bf16, q.lddvb = None, dvb.ptr, dvb.ld
    q.db = E._p(db)
    q.ws = None
    # padding-aware row pass: the last layer of a masked assignment stack gets its only upstream gradient from
    # dza = dT.Wp, whose pad rows are exactly zero (masked softmax backward): those rows are written as zeros unread
    if last and getattr(ctx, 'pad_grad_zero', False) and ctx.nb is not None and dxn is None and dout_ptr is None:
        q.nb_zero = E._p(ctx.nb)
    wsf = ws.f(int(load().gp_gcn_layer_bwd_ws_x(C.byref(q)))) if has_b else None
    q.ws = E._p(wsf)
    call('gp_gcn_layer_bwd_x', C.byref(q), st)
    # dW = U^T dV : U stored [rows, din] = M-major A ; dV N-major B ; split-K over the rows
    dw = ws.f(din, wc)
    tcgemm(Op(ub.ptr, ub.ld, 0), MN, Op(dvb.ptr, dvb.ld, 0), MN, din, wc, rows, 1, Cf=(dw.data_ptr(), wc, 0),
           split_k=pick_split(din, wc, rows))
    if has_b and wc != dout:
        db = db[:wc]
    return dvb, dw, db


def _layer_du(ws, ctx, l, dvb, dub):
    """dU = dV W^T -> bf16 `dub` ([B*N, din] view, possibly a column half): dV K-major; W stored [din, dout] = K-major B."""
    xb, din, dout, off, ub, y, rnorm, mean, invstd, wb = ctx.layers[l]
    rows = ctx.B * ctx.N
    tcgemm(Op(dvb.ptr, dvb.ld, 0), KM, Op(wb.ptr, wb.ld, 0), KM, rows, din, dout, 1, Cb=Op(dub.ptr, dub.ld, 0))


def stack_backward(ws, ctx, dz_ptr, lddz, dout_ptr, arg_ptr, ldo, need_dx, dadj):
    B, N = ctx.B, ctx.N
    L = len(ctx.layers)
    grads = [None] * L
    dxn = None
    nbp = E._p(ctx.nb)
    lim = int(ctx.nb is not None)
    for l in reversed(range(L)):
        xb, din = ctx.layers[l][0], ctx.layers[l][1]
        dvb, dw, db = _layer_backward_head(ws, ctx, l, dz_ptr, lddz, E._p(dxn), 0, dout_ptr, arg_ptr, ldo)
        grads[l] = (dw, db)
        need_dx_l = need_dx or l > 0
        dx = None
        if need_dx_l or dadj is not None:
            dub = bfbuf(ws, B, N, din)
            _layer_du(ws, ctx, l, dvb, dub)
            if need_dx_l:
                # dX = A^T dU : A stored [k rows, m cols] = M-major ; dU N-major
                dx = ws.f(B, N, din)
                tcgemm(ctx.adjb, MN, dub, MN, N, din, N, B, Cf=(dx.data_ptr(), din, N * din), lim=nbp, lim_m=lim,
                       lim_k=lim)
                if getattr(ctx, 'drops', None) is not None and ctx.drops[l] is not None:
                    pd, sd = ctx.drops[l]
                    E.dropout(dx.data_ptr(), din, B * N, din, pd, sd, dx.data_ptr(), din)
            if dadj is not None:
                # dA += dU X^T : dU K-major ; B[n=node, k=din] = X stored [node rows, din cols] = K-major
                tcgemm(dub, KM, xb, KM, N, N, din, B, Cf=(dadj.data_ptr(), dadj.shape[2], N * dadj.shape[2]), beta=1.0)
        dxn = dx
    return grads, dxn


def dual_stack_backward(ws, cE, cA, dzE_ptr, lddzE, doutE_ptr, argE_ptr, ldo, dzA_ptr, lddzA, dz_bf16=False):
    """Backward of dual_stack_forward (level 0: no dX of the first layer, no dA): per layer both tails, then ONE
    dX = A^T [dU_e | dU_a] pass over the adjacency for both stacks.  dz_bf16: both dz sources are bf16 buffers (lddz in
    elements); the lock-step dX is then kept in bf16 as well."""
    B, N = cE.B, cE.N
    xb16 = bool(dz_bf16)
    esz = 2 if xb16 else 4
    L = len(cE.layers)
    gE, gA = [None] * L, [None] * L
    nbp, lim = E._p(cE.nb), int(cE.nb is not None)
    dxcat, wcat, de_in, dx_ptr, dx_ld = None, 0, 0, None, 0
    for l in reversed(range(L)):
        dinE, dinA = cE.layers[l][1], cA.layers[l][1]
        if dxcat is None:
            xE = xA = None
        else:
            xE, xA = dx_ptr, dx_ptr + de_in * esz
        dvE, dw, db = _layer_backward_head(ws, cE, l, dzE_ptr, lddzE, xE, dx_ld, doutE_ptr, argE_ptr, ldo,
                                           dz_bf16=dz_bf16, dxn_bf16=xb16)
        gE[l] = (dw, db)
        dvA, dw, db = _layer_backward_head(ws, cA, l, dzA_ptr, lddzA, xA, dx_ld, None, None, 0,
                                           dz_bf16=dz_bf16, dxn_bf16=xb16)
        gA[l] = (dw, db)
        if l > 0:
            wcat, de_in = dinE + dinA, dinE
            ducat = bfbuf(ws, B, N, wcat)
            _layer_du(ws, cE, l, dvE, Op(ducat.ptr, ducat.ld, 0))
            _layer_du(ws, cA, l, dvA, Op(ducat.ptr + dinE * 2, ducat.ld, 0))
            if xb16:
                dxcat = bfbuf(ws, B, N, wcat)
                dx_ptr, dx_ld = dxcat.ptr, dxcat.ld
                tcgemm(cE.adjb, MN, ducat, MN, N, wcat, N, B, Cb=dxcat, lim=nbp, lim_m=lim, lim_k=lim)
            else:
                dxcat = ws.f(B, N, wcat)
                dx_ptr, dx_ld = dxcat.data_ptr(), wcat
                tcgemm(cE.adjb, MN, ducat, MN, N, wcat, N, B, Cf=(dx_ptr, wcat, N * wcat), lim=nbp, lim_m=lim,
                       lim_k=lim)
    return gE, gA


# ------------------------------------------------------------------------------------------
def softmax_forward(ws, S, nb, B, N, K):
    """In-place masked softmax (encoders.py:1273-1275) that also emits the bf16 operand copy of S."""
    sb = bfbuf(ws, B, N, K)
    if K % 4 == 0 and K <= 2048:
        call('gp_softmax_mask_fwd_x', S.data_ptr(), E._p(nb), B, N, K, sb.ptr, sb.ld, E._stream())
    else:
        call('gp_softmax_mask_fwd', S.data_ptr(), E._p(nb), B, N, K, E._stream())
        cvt(ws, S.data_ptr(), K, B * N, K, out=Op(sb.ptr, sb.ld, 0))
    return sb


CHAIN_POOLING = None     # None: decided by GP_CHAIN; True / False: set by the caller (tests, bench)


def bf16_grads():
    """Gradient intermediates between GEMM epilogues and the layer-backward row kernels (dZ of the pooling backward, dza
    of the assignment head, the lock-step dX) are kept in bf16 -- like dV always was: half the bytes to write and to read
    (-3.5 GB per cfg4 step).  GP_F32_GRADS=1 keeps them in fp32."""
    return not os.environ.get('GP_F32_GRADS')


def stack_takes_bf16_grads(ctx):
    """Every layer of the stack is served by a layer-backward kernel instantiated for bf16 dz / dxn (elsewhere bf16
    sources go through run-time branches and are slower than fp32 ones: the caller then keeps fp32 intermediates)."""
    L = len(ctx.layers)
    lib = load()
    return ctx.aligned and all(lib.gp_gcn_layer_bwd_bf16_sources_fast(ctx.B, ctx.douts[l], int(bool(ctx.bn and l < L - 1)))
                               for l in range(L))


def chain_ok(sb, adjb, N, K):
    """The chained S^T A S kernel (gp_pool_chain_bf16) takes dense per-graph operands and at most 512 clusters.  It is
    OPT-IN (GP_CHAIN=1 or engine_tc.CHAIN_POOLING = True): measured on B200 at cfg4 (K = 512) it is slower than T = S^T A
    on CTA pairs followed by A' = T S (1.5 vs 1.0 ms), although it moves 1 GB less through HBM."""
    want = CHAIN_POOLING if CHAIN_POOLING is not None else bool(os.environ.get('GP_CHAIN'))
    return want and K <= 512 and sb.sb == N * sb.ld and adjb.sb == N * adjb.ld


def pool_forward(ws, sb, zb, adjb, nb, B, N, K, Fw, keep_t=True):
    """X' = S^T Z, A' = S^T A S (encoders.py:1278-1279).  Default: T = S^T A and A' = T S as two launches (the first
    on cta_group::2 CTA pairs).  With the chain selected (chain_ok) and K <= 512, A' comes from ONE launch that keeps T on
    chip (TMEM -> bf16 shared-memory tile -> second tcgen05.mma); T is written to HBM (bf16, once) only when the backward
    will need it (`keep_t`)."""
    nbp, lim = E._p(nb), int(nb is not None)
    xp, xpb = ws.f(B, K, Fw), bfbuf(ws, B, K, Fw)
    tcgemm(sb, MN, zb, MN, K, Fw, N, B, Cf=(xp.data_ptr(), Fw, K * Fw), Cb=xpb, lim=nbp, lim_k=lim)
    ap, apb = ws.f(B, K, K), bfbuf(ws, B, K, K)
    if chain_ok(sb, adjb, N, K):
        tb = bfbuf(ws, B, K, N) if keep_t else None
        ent = _ORDER.get(nbp) if nbp is not None else None
        order = ent[0].data_ptr() if (ent is not None and ent[1] == B) else None
        call('gp_pool_chain_bf16', sb.ptr, C.c_longlong(sb.ld), adjb.ptr, C.c_longlong(adjb.ld), nbp, order, B, N, K,
             None if tb is None else tb.ptr, C.c_longlong(0 if tb is None else tb.ld), ap.data_ptr(), C.c_longlong(K),
             apb.ptr, C.c_longlong(apb.ld), E._stream())
        return sb, xp, xpb, tb, ap, apb
    tb = bfbuf(ws, B, K, N)
    tcgemm(sb, MN, adjb, MN, K, N, N, B, Cb=tb, lim=nbp, lim_k=lim, lim_n=lim)
    tcgemm(tb, KM, sb, MN, K, K, N, B, Cf=(ap.data_ptr(), K, K * K), Cb=apb, lim=nbp, lim_k=lim)
    return sb, xp, xpb, tb, ap, apb


def pool_backward(ws, dxp, dap, sb, zb, adjb, tb, nb, B, N, K, Fw, ds, acc_ds, dadj, asym=None, dz_bf16=False):
    """asym: device flag from adj_prepare (0 = every adjacency of the batch is symmetric) or None.  For a symmetric
    A, T^T = A S, so T^T dA' + A (S dA'^T) = T^T (dA' + dA'^T): the N x N x K product and S dA'^T are skipped
    on the device (no host sync): the kernels read the flag."""
    nbp, lim = E._p(nb), int(nb is not None)
    dxpb = cvt(ws, dxp.data_ptr(), Fw, B * K, Fw, B=B)
    ldap = dap.shape[2]                                  # dA' rows may be padded (see _bwd_tc)
    dapb = cvt(ws, dap.data_ptr(), ldap, B * K, K, B=B)
    if dz_bf16:                                          # dZ = S dX' as a bf16 buffer (an Op; consumed by gp_gcn_layer_bwd_x)
        dz = bfbuf(ws, B, N, Fw)
        tcgemm(sb, KM, dxpb, MN, N, Fw, K, B, Cb=dz, lim=nbp, lim_m=lim)
    else:
        dz = ws.f(B, N, Fw)
        tcgemm(sb, KM, dxpb, MN, N, Fw, K, B, Cf=(dz.data_ptr(), Fw, N * Fw), lim=nbp, lim_m=lim)
    dsf = (ds.data_ptr(), K, N * K)
    wsb = bfbuf(ws, B, N, K)
    cond = E._p(asym)
    tcgemm(sb, KM, dapb, KM, N, K, K, B, Cb=wsb, lim=nbp, lim_m=lim, cond=cond, cond_npairs=0)
    dapx = dapb
    if asym is not None:                                 # dA' + dA'^T when symmetric, dA' otherwise
        dapx = bfbuf(ws, B, K, K)
        call('gp_sym_select_bf16', dap.data_ptr(), C.c_longlong(ldap), B, K, cond, dapx.ptr, dapx.ld, E._stream())
    # dS (+)= Z dX'^T + T^T dA' + A (S dA'^T): three products accumulated in TMEM, one pass over dS
    tcgemm_multi([(zb, KM, dxpb, KM, Fw, 0), (tb, MN, dapx, MN, K, 0), (adjb, KM, wsb, MN, N, lim)], N, K, B,
                 Cf=dsf, beta=1.0 if acc_ds else 0.0, lim=nbp, lim_m=lim, cond=cond, cond_npairs=2)
    if dadj is not None:
        w2b = bfbuf(ws, B, N, K)
        tcgemm(sb, KM, dapb, MN, N, K, K, B, Cb=w2b)
        tcgemm(w2b, KM, sb, KM, N, N, K, B, Cf=(dadj.data_ptr(), dadj.shape[2], N * dadj.shape[2]), beta=1.0)
    return dz


def assign_linear_fwd(ws, zab, Fa, rows, wp, bp, Kp=0):
    """T = za.Wp^T + bp.  Kp > K (padded cluster count): Fa is the padded concat width; the extra rows / columns of
    the bf16 weight are zero and the extra logits get a bias of -1e30, i.e. probability exactly 0 after the softmax
    (dead clusters: S, X' = S^T Z and A' = S^T A S only gain zero columns / rows)."""
    K, Fr = int(wp.shape[0]), int(wp.shape[1])
    if not Kp or Kp == K:
        wpb = cvt(ws, wp.data_ptr(), Fa, K, Fa)                                         # [K, r8(Fa)]
        T = ws.f(rows, K)
        tcgemm(Op(zab.ptr, zab.ld, 0), KM, Op(wpb.ptr, wpb.ld, 0), KM, rows, K, Fa, 1, Cf=(T.data_ptr(), K, 0),
               bias=E._p(bp))
        return T, wpb
    st = E._stream()
    wpb = bfbuf(ws, 1, Kp, Fa)
    call('gp_fill_f32', wpb.ptr, C.c_longlong(Kp * wpb.ld // 2), C.c_float(0.0), st)    # bf16 zeros, two per float
    cvt(ws, wp.data_ptr(), Fr, K, Fr, out=wpb)
    bpad = ws.f(Kp)
    if bp is None:
        call('gp_fill_f32', bpad.data_ptr(), C.c_longlong(K), C.c_float(0.0), st)
        call('gp_fill_f32', bpad.data_ptr() + 4 * K, C.c_longlong(Kp - K), C.c_float(-1e30), st)
    else:
        call('gp_pad_copy_f32', bp.data_ptr(), C.c_longlong(K), C.c_longlong(1), K, bpad.data_ptr(),
             C.c_longlong(Kp), C.c_longlong(1), Kp, C.c_float(-1e30), st)
    T = ws.f(rows, Kp)
    tcgemm(Op(zab.ptr, zab.ld, 0), KM, Op(wpb.ptr, wpb.ld, 0), KM, rows, Kp, Fa, 1, Cf=(T.data_ptr(), Kp, 0),
           bias=bpad.data_ptr())
    return T, wpb


def assign_head_bwd(ws, S, ds, nb, B, N, zab, Fa, wpb, K, has_bias, Kreal=0, Fa_real=0, dza_bf16=False):
    """Backward of S = softmax(assign_pred(za)) * mask: dT (bf16 operand + bias gradient in one pass), then
    dWp = dT^T za (split-K) and dza = dT Wp.  K / Fa may be the padded widths; the parameter gradients are produced
    at the real ones (Kreal x Fa_real: the real clusters / concat columns come first)."""
    Kreal, Fa_real = Kreal or K, Fa_real or Fa
    st = E._stream()
    rows = B * N
    dtb = bfbuf(ws, 1, rows, K)
    dbp = ws.f(K) if has_bias else None
    if K % 4 == 0 and K <= 2048:
        wsb = ws.f((148 * 16 + 256) * K) if has_bias else None
        call('gp_softmax_mask_bwd_x', S.data_ptr(), ds.data_ptr(), E._p(nb), B, N, K, None, dtb.ptr, dtb.ld,
             E._p(dbp), E._p(wsb), st)
    else:
        dt = ws.f(B, N, K)
        call('gp_softmax_mask_bwd', S.data_ptr(), ds.data_ptr(), E._p(nb), B, N, K, dt.data_ptr(), st)
        cvt(ws, dt.data_ptr(), K, rows, K, out=dtb)
        if has_bias:
            cs = ws.f(256 * K)
            call('gp_colsum_f32', E._p(dt), C.c_longlong(rows), K, C.c_longlong(K), E._p(dbp), 0, E._p(cs), st)
    dwp = ws.f(Kreal, Fa_real)
    tcgemm(Op(dtb.ptr, dtb.ld, 0), MN, Op(zab.ptr, zab.ld, 0), MN, Kreal, Fa_real, rows, 1,
           Cf=(dwp.data_ptr(), Fa_real, 0), split_k=pick_split(Kreal, Fa_real, rows))
    if has_bias and Kreal != K:
        dbp = dbp[:Kreal]
    if dza_bf16:                                         # bf16 buffer (an Op with ld = r8(Fa))
        dza = bfbuf(ws, 1, rows, Fa)
        tcgemm(Op(dtb.ptr, dtb.ld, 0), KM, Op(wpb.ptr, wpb.ld, 0), MN, rows, Fa, K, 1, Cb=Op(dza.ptr, dza.ld, 0))
    else:
        dza = ws.f(rows, Fa)
        tcgemm(Op(dtb.ptr, dtb.ld, 0), KM, Op(wpb.ptr, wpb.ld, 0), MN, rows, Fa, K, 1, Cf=(dza.data_ptr(), Fa, 0))
    return dwp, dbp, dza


def upper_band_ok(sb, adjb, N, mode, adj_flags):
    """gp_linkloss_tc mode 2 / gp_gemm_bf16x.tri: for a symmetric {0,1} adjacency (decided on the device from the
    gp_adj_prepare flags) G = dl/dP is written as its upper diagonal band only and the backward reads that band twice
    (once transposed).  Needs the BCE loss, the flags and 32-byte aligned rows of at least round_up(N, 32) elements
    (adjacency operands and G are allocated that way: bfbuf(..., pad=32))."""
    return (mode == 0 and adj_flags is not None and adjb.ld % 16 == 0 and adjb.ld >= (N + 31) // 32 * 32 and adjb.ptr % 32 == 0 and
            not os.environ.get('GP_NO_UPPER_G'))


def linkloss_forward(ws, sb, adjb, nb, B, N, K, need_grad, mode=0, adj_flags=None):
    """Fused tensor-core link loss: P = S S^T tiles stay in TMEM, the epilogue does the masked BCE
    against the bf16 adjacency and writes gsym (bf16).  Returns (partial, n_partial, gsym op, upper)."""
    nbp = E._p(nb)
    npart = int(load().gp_linkloss_tc_partials(B, N))
    partial = ws.f(npart + 256)                      # +256: scratch of the two-stage finalisation
    gs = bfbuf(ws, B, N, N, pad=32) if need_grad else None
    upper = upper_band_ok(sb, adjb, N, mode, adj_flags) and (gs is None or (gs.ld % 16 == 0 and gs.ptr % 32 == 0 and
                                                                      gs.ld >= (N + 31) // 32 * 32))
    call('gp_linkloss_tc', sb.ptr, sb.ld, adjb.ptr, adjb.ld, nbp, B, N, K, partial.data_ptr(),
         None if gs is None else gs.ptr, N if gs is None else gs.ld, 2 if upper else mode, E._p(adj_flags), E._stream())
    return partial, npart, gs, upper


def linkloss_backward(ws, gs, sb, nb, B, N, K, inv, g_ptr, dS=None, asym=None, upper=False):
    nbp, lim = E._p(nb), int(nb is not None)
    if dS is None:
        dS = ws.f(B, N, K)
    # dS = (G + G^T).S : G K-major, then the same buffer read M-major (= G^T), accumulated.  A symmetric adjacency
    # makes G symmetric (P = S S^T is): the kernel then runs ONE full contraction with alpha doubled -- the first
    # product only, or (`upper`: G holds its upper diagonal band only) the band read directly for k >= the row
    # block's diagonal and transposed for k below it (gp_gemm_bf16x.tri).
    cf = (dS.data_ptr(), K, N * K)
    tcgemm_multi([(gs, KM, sb, MN, N, lim), (gs, MN, sb, MN, N, lim)], N, K, B, Cf=cf, alpha=inv, alpha_dev=g_ptr,
                 lim=nbp, lim_m=lim, cond=E._p(asym), cond_npairs=2 if upper else 1, cond_alpha=2.0,
                 tri=1 if upper else 0)
    return dS
