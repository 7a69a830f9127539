"""Drop-in replacement for the DiffPool half of the reference's ``encoders`` module.

Same class names, constructor signatures, ``forward(x, adj, batch_num_nodes, assign_x=...)``,
``loss(...)``, attributes (``assign_tensor``, ``link_loss``) and state-dict keys as
``/root/reference/encoders.py:976-1334`` so that ``train.py`` / ``cross_val.py`` can run against
it unchanged (``import encoders`` -> this module, see graph_pooling_b200/shim.py).

Every tensor op of forward / loss / backward runs in hand-written sm_100a kernels behind the C
ABI (include/gp_b200.h); see engine.py for the schedule.  Deviations from the shipped reference
text are the documented repairs R1-R11 (oracle/diffpool_oracle.py header).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import engine as E
from . import engine_pk as PK
from . import engine_tc as T
from ._lib import call

__all__ = ['GraphConv', 'GcnEncoderGraph', 'GcnSet2SetEncoder', 'Set2Set', 'SoftPoolingGcnEncoder']

# precision new encoders start with: 0 = fp32 FFMA (parity anchor), 1 = bf16 tensor cores; `model.precision` overrides
DEFAULT_PRECISION = E.F32


# ------------------------------------------------------------------------------------------
# autograd glue: one Function for the whole encoder, one for the loss
# ------------------------------------------------------------------------------------------
class _Plan:
    """Static description of one forward call (indices into the flat parameter list)."""
    pass


def _wb(params, pair):
    iw, ib = pair
    return params[iw], (None if ib is None else params[ib])


def _kpad(K):
    """Cluster count the tensor-core schedule runs a pooling level at: K rounded up to a multiple of 8 so that every
    row of S / the assignment concat / A' is 16-byte aligned (TMA and vector paths).  The extra clusters are DEAD: zero
    weights, logit bias -1e30 => probability exactly 0, zero rows / columns of X' and A', excluded from the next
    level's readout and masked like pad nodes in its assignment.  GP_NO_KPAD=1 disables it (debug)."""
    return K if os.environ.get('GP_NO_KPAD') else T.r8(K)


def _fwd_tc(ctx, plan, x, adj, assign_x, params):
    """GP_BF16 forward: same schedule as _EncoderFn.forward with the contractions on tcgen05."""
    st = E._stream()
    ws = E.Workspace(x.device)
    B, N, D = x.shape
    nb = plan.nb_dev
    conv = lambda pairs: ([_wb(params, p)[0] for p in pairs], [_wb(params, p)[1] for p in pairs])
    P, Fw = plan.num_pooling, plan.F
    ldo = Fw * (P + 1)
    out, arg = ws.f(B, ldo), ws.i(B, ldo)
    if isinstance(adj, T.PreparedAdjacency):
        adjb, aflags = adj.op, adj.flags
    else:
        adjb, aflags = T.adj_prepare(ws, adj, nb, B, N)   # bf16 operand + [not symmetric, not {0,1}] device flags
    if x.dtype == torch.bfloat16:          # features already in the operand precision (compact feed): used as they are
        xb = T.Op(x.data_ptr(), D, N * D, x)
    else:
        xb = T.cvt(ws, x.data_ptr(), D, B * N, D, B=B)
    w0, b0 = conv(plan.emb)
    xab, xa_d, pre_as = xb, D, None
    if plan.soft and assign_x is not x:
        xa_d = assign_x.shape[2]
        xab = T.cvt(ws, assign_x.data_ptr(), xa_d, B * N, xa_d, B=B)
    dual = False
    if plan.soft:
        wa0, ba0 = conv(plan.assign[0])
        dual = T.dual_ok(w0, wa0) and not os.environ.get('GP_NO_DUAL') and not any(plan.emb_drop)
    if dual:        # embedding + level-0 assignment GCN in lock-step: one pass over A per layer for both
        K0 = plan.assign_dims[0]
        (z, zb, c_emb), pre_as = T.dual_stack_forward(ws, xb, D, xab, xa_d, adjb, nb, B, N, w0, b0, plan.bn, wa0, ba0,
                                                      True, pad_lastA=_kpad(K0) if _kpad(K0) != K0 else 0)
    else:
        z, zb, c_emb = T.stack_forward(ws, xb, D, adjb, nb, B, N, w0, b0, plan.bn, drops=plan.emb_drop, seed=plan.seed)
    call('gp_readout_max_fwd', z.data_ptr(), Fw, E._p(nb) if plan.soft else None, B, N, Fw,
         out.data_ptr(), arg.data_ptr(), ldo, st)
    levels, S0 = [], None
    plan.adjb, plan.sb0, plan.asym, plan.adj_flags, plan.S0_full = adjb, None, aflags[0:1], aflags, None
    vis_S = []
    if plan.soft:
        # cur_N rows are allocated per graph at this level, of which cur_Nr exist (cur_N > cur_Nr: dead clusters of a
        # padded level above, masked through cur_nb like pad nodes)
        cur_adjb, cur_nb, cur_N, cur_Nr, cur_zb = adjb, nb, N, N, zb
        for i in range(P):
            Kr = plan.assign_dims[i]
            K = _kpad(Kr)                                # width this level runs at (dead clusters beyond Kr)
            pad = K if K != Kr else 0
            wa, ba = conv(plan.assign[i])
            if i == 0 and pre_as is not None:
                za, zab, c_as = pre_as
            else:
                # level 0 with assign_x == x: both GCNs start from the same U = A.x -- compute it once
                u0 = c_emb.layers[0][4] if (i == 0 and xab is xb) else None
                za, zab, c_as = T.stack_forward(ws, xab, xa_d, cur_adjb, cur_nb, B, cur_N, wa, ba, True, u0=u0,
                                                pad_last=pad, h32=False)
            c_as.pad_grad_zero = True                   # S's pad rows are masked: no gradient reaches this stack's pad rows
            Fa = za.shape[2]
            wp, bp = _wb(params, plan.assign_pred[i])
            Tl, wpb = T.assign_linear_fwd(ws, zab, Fa, B * cur_N, wp, bp, Kp=pad)
            S = Tl.view(B, cur_N, K)
            sb = T.softmax_forward(ws, S, cur_nb, B, cur_N, K)
            sb, xp, xpb, tb, ap, apb = T.pool_forward(ws, sb, cur_zb, cur_adjb, cur_nb, B, cur_N, K, Fw,
                                                      keep_t=any(ctx.needs_input_grad))
            wq, bq = conv(plan.post[i])
            nbk = None
            if pad:                                      # "node counts" of the pooled level: Kr real clusters of K
                nbk = ws.i(B)
                call('gp_fill_i32', nbk.data_ptr(), C.c_longlong(B), Kr, st)
            z2, z2b, c_post = T.stack_forward(ws, xpb, Fw, apb, nbk, B, K, wq, bq, plan.bn_post,
                                              drops=plan.post_drop[i], seed=plan.seed + 4096 * (i + 1))
            # the reference's max over the pooled level's clusters (encoders.py:1287): the Kr real rows only
            call('gp_readout_max_fwd_x', z2.data_ptr(), Fw, K, None, B, Kr, Fw, out.data_ptr() + (i + 1) * Fw * 4,
                 arg.data_ptr() + (i + 1) * Fw * 4, ldo, st)
            levels.append(dict(K=K, Kr=Kr, N=cur_N, nb=cur_nb, adjb=cur_adjb, zb=cur_zb, S=S.detach(), sb=sb, zab=zab,
                               Fa=Fa, Fa_r=int(wp.shape[1]), c_as=c_as, tb=tb, c_post=c_post, wpb=wpb,
                               has_bp=bp is not None, asym=plan.asym if i == 0 else None))
            Sv = S if (K == Kr and cur_N == cur_Nr) else S[:, :cur_Nr, :Kr]     # what the caller sees: real rows / clusters
            vis_S.append(Sv)
            if i == 0:
                S0, plan.sb0 = Sv, sb
                plan.S0_full = S.detach() if pad else None
            xab, xa_d = xpb, Fw
            cur_adjb, cur_nb, cur_N, cur_Nr, cur_zb = apb, nbk, K, Kr, z2b
    lin = [_wb(params, p) for p in plan.pred]
    ypred, acts = E.mlp_fwd(ws, out.data_ptr(), ldo, B, lin)
    ctx.tape = dict(plan=plan, params=params, B=B, N=N, emb=c_emb, levels=levels, out=out, arg=arg, ldo=ldo,
                    acts=acts, lin=lin, x=x, adj=adj, dual=dual)
    if plan.soft:
        # detached aliases: the tape holds `plan` and the returned S0 gets this node as grad_fn; storing S0
        # itself would close a reference cycle through the autograd node that only backward() breaks
        plan.all_S = [v.detach() for v in vis_S]
        return ypred, S0
    return ypred


def _padded_grad(ws, plan, dS0, B, N, Kr, K):
    """Gradient of the visible S0 [B,N,Kr] as a [B,N,K] buffer (K = padded width).  _LossFn hands back a narrowed view
    of its own padded buffer (tagged in plan.ds_tag): that buffer is used in place; anything else is copied."""
    if K == Kr:
        return E._chk(dS0, 'grad of assign_tensor')
    if dS0.is_cuda and dS0.dtype == torch.float32 and dS0.data_ptr() == getattr(plan, 'ds_tag', None) \
            and tuple(dS0.stride()) == (N * K, K, 1):
        return torch.as_strided(dS0, (B, N, K), (N * K, K, 1))
    src = E._chk(dS0, 'grad of assign_tensor')
    out = ws.f(B, N, K)
    call('gp_pad_copy_f32', src.data_ptr(), C.c_longlong(Kr), C.c_longlong(B * N), Kr, out.data_ptr(),
         C.c_longlong(K), C.c_longlong(B * N), K, C.c_float(0.0), E._stream())
    return out


def _bwd_tc(ctx, tape, dypred, dS0):
    plan, params = tape['plan'], tape['params']
    B = tape['B']
    st = E._stream()
    ws = E.Workspace(tape['x'].device)
    Fw, P, ldo = plan.F, plan.num_pooling, tape['ldo']
    grads = [None] * len(params)

    def put(pairs, gl):
        for (iw, ib), (dw, db) in zip(pairs, gl):
            grads[iw] = dw
            if ib is not None:
                grads[ib] = db

    if dypred is None:
        dypred = ws.z(B, plan.label_dim)
    dypred = E._chk(dypred, 'grad of ypred')
    dout = ws.f(B, ldo)
    put(plan.pred, E.mlp_bwd(ws, dypred, B, tape['acts'], tape['lin'], dout.data_ptr(), ldo))
    dout_p, arg_p = dout.data_ptr(), tape['arg'].data_ptr()
    dz_dense = None
    if plan.soft:
        levels = tape['levels']
        # dA' accumulators: row stride padded to 4 floats so the read-modify-write GEMM epilogues stay on their
        # 16-byte vector path for any K (K = 250 / 1250 at the DD / ragged configs)
        d_ap = [ws.z(B, lv['K'], (lv['K'] + 3) & ~3) for lv in levels]
        dxp_extra = [None] * P
        dz_next = None
        for i in reversed(range(P)):
            lv = levels[i]
            K, Ni = lv['K'], lv['N']
            gl, dxp = T.stack_backward(ws, lv['c_post'], None if dz_next is None else dz_next.data_ptr(), Fw,
                                       dout_p + (i + 1) * Fw * 4, arg_p + (i + 1) * Fw * 4, ldo, True, d_ap[i])
            put(plan.post[i], gl)
            if dxp_extra[i] is not None:
                call('gp_axpy_f32', dxp_extra[i].data_ptr(), dxp.data_ptr(), C.c_longlong(dxp.numel()),
                     C.c_float(1.0), st)
            if i == 0 and dS0 is not None:
                ds, acc_ds = _padded_grad(ws, plan, dS0, B, Ni, lv['Kr'], K), 1
            else:
                ds, acc_ds = ws.f(B, Ni, K), 0
            # level 0 with lock-step stacks: dZ, dza and the per-layer dX stay bf16 between the GEMM epilogues and the
            # layer-backward row kernels (needs 16-byte aligned concat slots: Fw % 8 == 0 and Fa % 8 == 0)
            g16 = bool(i == 0 and tape['dual'] and T.bf16_grads() and Fw % 8 == 0 and lv['Fa'] % 8 == 0 and
                       T.stack_takes_bf16_grads(tape['emb']) and T.stack_takes_bf16_grads(lv['c_as']))
            dz = T.pool_backward(ws, dxp, d_ap[i], lv['sb'], lv['zb'], lv['adjb'], lv['tb'], lv['nb'], B, Ni, K, Fw,
                                 ds, acc_ds, None if i == 0 else d_ap[i - 1], asym=lv['asym'], dz_bf16=g16)
            dwp, dbp, dza = T.assign_head_bwd(ws, lv['S'], ds, lv['nb'], B, Ni, lv['zab'], lv['Fa'], lv['wpb'], K,
                                              lv['has_bp'], Kreal=lv['Kr'], Fa_real=lv['Fa_r'], dza_bf16=g16)
            iw, ib = plan.assign_pred[i]
            grads[iw] = dwp
            if ib is not None:
                grads[ib] = dbp
            if i == 0 and tape['dual']:
                # embedding + assignment GCN backward in lock-step (one A^T.dU pass per layer for both)
                if g16:
                    gE, gA = T.dual_stack_backward(ws, tape['emb'], lv['c_as'], dz.ptr, dz.ld, dout_p, arg_p, ldo,
                                                   dza.ptr, dza.ld, dz_bf16=True)
                else:
                    gE, gA = T.dual_stack_backward(ws, tape['emb'], lv['c_as'], dz.data_ptr(), Fw, dout_p, arg_p, ldo,
                                                   dza.data_ptr(), lv['Fa'])
                put(plan.emb, gE)
                put(plan.assign[0], gA)
                return (None, None, None, None, None) + tuple(_deliver(plan, params, grads))
            gl, dxa = T.stack_backward(ws, lv['c_as'], dza.data_ptr(), lv['Fa'], None, None, 0, i > 0,
                                       None if i == 0 else d_ap[i - 1])
            put(plan.assign[i], gl)
            if i > 0:
                dxp_extra[i - 1] = dxa
            dz_next = dz
        dz_dense = dz_next
    gl, _ = T.stack_backward(ws, tape['emb'], None if dz_dense is None else dz_dense.data_ptr(), Fw, dout_p, arg_p,
                             ldo, False, None)
    put(plan.emb, gl)
    return (None, None, None, None, None) + tuple(_deliver(plan, params, grads))


def _deliver(plan, params, grads):
    """Hand the parameter gradients of a backward pass to autograd -- or, in sink mode (_apply_encoder), add them
    into the attached dp.FlatGradients buffer with one gp_multi_axpy_f32 launch; autograd then sees no parameter at all."""
    real = getattr(plan, 'sink_params', None)
    if real is None:
        return grads
    left = plan.grad_sink.accumulate(real, grads)
    for p, g in zip(real, left):               # a parameter outside the buffer (not expected): plain accumulation
        if g is not None:
            p.grad = g if p.grad is None else p.grad + g
    return [None] * len(grads)


def _apply_encoder(plan, x, adj, x_a, params):
    """_EncoderFn.apply.  With a gradient sink attached (dp.FlatGradients.attach) the parameters enter the autograd
    node DETACHED and a fresh empty `anchor` tensor carries requires_grad instead: the backward then delivers the
    parameter gradients itself (one launch) and autograd runs no AccumulateGrad node for them -- an AccumulateGrad
    node that receives an undefined gradient would still register its own (creation-time) stream for the engine's
    final synchronisation, which breaks CUDA-graph capture of the step."""
    if plan.grad_sink is not None and torch.is_grad_enabled() and any(p.requires_grad for p in params):
        plan.sink_params = list(params)
        anchor = torch.empty(0, device=x.device, requires_grad=True)
        return _EncoderFn.apply(plan, x, adj, x_a, anchor, *[p.detach() for p in params])
    plan.sink_params = None
    return _EncoderFn.apply(plan, x, adj, x_a, None, *params)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, adj, assign_x, anchor, *params):
        ctx.set_materialize_grads(False)
        if plan.precision == T.BF16:
            return _fwd_tc(ctx, plan, x, adj, assign_x, params)
        if getattr(plan, 'packed', False):            # ENZYMES-sized graphs: packed rows, one launch per phase
            ypred, S0, tape = PK.forward(plan, x, adj, assign_x, params, _wb)
            tape.update(plan=plan, params=params, packed=True)
            ctx.tape = tape
            plan.all_S = [S0.detach()]
            return ypred, S0
        st = E._stream()
        ws = E.Workspace(x.device)
        B, N, D = x.shape
        prec = plan.precision
        nb = plan.nb_dev
        tape = {'plan': plan, 'params': params, 'B': B, 'N': N}
        conv = lambda pairs: ([_wb(params, p)[0] for p in pairs], [_wb(params, p)[1] for p in pairs])
        P = plan.num_pooling
        Fw = plan.F
        out = ws.f(B, Fw * (P + 1))
        arg = ws.i(B, Fw * (P + 1))
        ldo = Fw * (P + 1)

        w0, b0 = conv(plan.emb)
        z, c_emb = E.stack_forward(ws, x.data_ptr(), D, D, adj, nb, B, N, w0, b0, plan.add_self, plan.bn, prec,
                                   drops=plan.emb_drop, seed=plan.seed)
        # base path: no mask ever (encoders.py:1087 builds it, nothing uses it); soft: mask of :1078
        call('gp_readout_max_fwd', z.data_ptr(), Fw, E._p(nb) if plan.soft else None, B, N, Fw,
             out.data_ptr(), arg.data_ptr(), ldo, st)
        tape['emb'] = c_emb
        levels = []
        S0 = None
        if plan.soft:
            xa_ptr, xa_d = assign_x.data_ptr(), assign_x.shape[2]
            cur_adj, cur_nb, cur_N, cur_z = adj, nb, N, z
            for i in range(P):
                K = plan.assign_dims[i]
                wa, ba = conv(plan.assign[i])
                za, c_as = E.stack_forward(ws, xa_ptr, xa_d, xa_d, cur_adj, cur_nb, B, cur_N, wa, ba,
                                           plan.add_self, True, prec)
                Fa = za.shape[2]
                wp, bp = _wb(params, plan.assign_pred[i])
                # S = softmax(Linear(za)) * mask   (pad rows -> 0; masking za first is a no-op for real rows)
                S = E.linear_fwd(ws, za.data_ptr(), Fa, B * cur_N, wp, bp, relu=False).view(B, cur_N, K)
                call('gp_softmax_mask_fwd', S.data_ptr(), E._p(cur_nb), B, cur_N, K, st)
                xp, t, ap = ws.f(B, K, Fw), ws.f(B, K, cur_N), ws.f(B, K, K)
                call('gp_pool_fwd', S.data_ptr(), cur_z.data_ptr(), Fw, cur_adj.data_ptr(), E._p(cur_nb), B, cur_N,
                     K, Fw, xp.data_ptr(), t.data_ptr(), ap.data_ptr(), prec, st)
                wq, bq = conv(plan.post[i])
                z2, c_post = E.stack_forward(ws, xp.data_ptr(), Fw, Fw, ap, None, B, K, wq, bq, plan.add_self,
                                             plan.bn_post, prec, drops=plan.post_drop[i], seed=plan.seed + 4096 * (i + 1))
                call('gp_readout_max_fwd', z2.data_ptr(), Fw, None, B, K, Fw, out.data_ptr() + (i + 1) * Fw * 4,
                     arg.data_ptr() + (i + 1) * Fw * 4, ldo, st)
                levels.append(dict(K=K, N=cur_N, nb=cur_nb, adj=cur_adj, z=cur_z, S=S.detach(), za=za, Fa=Fa, c_as=c_as,
                                   xp=xp, t=t, ap=ap, c_post=c_post, wp=wp, has_bp=bp is not None))
                if i == 0:
                    S0 = S
                xa_ptr, xa_d = xp.data_ptr(), Fw
                cur_adj, cur_nb, cur_N, cur_z = ap, None, K, z2
        # concat=False (base only): the last layer's readout alone feeds pred_model (:1119)
        pin_off = 0 if plan.concat else Fw - plan.douts_last
        lin = [_wb(params, p) for p in plan.pred]
        ypred, acts = E.mlp_fwd(ws, out.data_ptr() + pin_off * 4, ldo, B, lin)
        tape.update(levels=levels, out=out, arg=arg, ldo=ldo, acts=acts, lin=lin, x=x, adj=adj, assign_x=assign_x)
        ctx.tape = tape
        if plan.soft:
            # detached aliases: the tape holds `plan` and the returned S0 gets this node as grad_fn; storing S0
            # itself would close a reference cycle through the autograd node that only backward() breaks
            plan.all_S = [lv['S'].detach() for lv in levels]
            return ypred, S0
        return ypred

    @staticmethod
    @once_differentiable
    def backward(ctx, dypred, dS0=None):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError('gp_b200: backward called twice (buffers were freed)')
        ctx.tape = None
        if tape['plan'].precision == T.BF16:
            return _bwd_tc(ctx, tape, dypred, dS0)
        plan, params = tape['plan'], tape['params']
        if tape.get('packed'):
            grads = PK.backward(plan, tape, params, dypred, dS0)
            return (None, None, None, None, None) + tuple(_deliver(plan, params, grads))
        B, N = tape['B'], tape['N']
        st = E._stream()
        ws = E.Workspace(tape['x'].device)
        prec = plan.precision
        Fw = plan.F
        P = plan.num_pooling
        grads = [None] * len(params)

        def put(pairs, gl):
            for (iw, ib), (dw, db) in zip(pairs, gl):
                grads[iw] = dw
                if ib is not None:
                    grads[ib] = db

        ldo = tape['ldo']
        if dypred is None:
            dypred = ws.z(B, plan.label_dim)
        dypred = E._chk(dypred, 'grad of ypred')
        pin_off = 0 if plan.concat else Fw - plan.douts_last
        dout = ws.f(B, ldo) if plan.concat else ws.z(B, ldo)
        gl = E.mlp_bwd(ws, dypred, B, tape['acts'], tape['lin'], dout.data_ptr() + pin_off * 4, ldo)
        put(plan.pred, gl)
        dout_p, arg_p = dout.data_ptr(), tape['arg'].data_ptr()

        dz_dense = None          # dense gradient of the current level's embedding concat buffer
        if plan.soft:
            levels = tape['levels']
            d_ap = [ws.z(B, lv['K'], lv['K']) for lv in levels]
            dxp_extra = [None] * P
            dz_next = None
            for i in reversed(range(P)):
                lv = levels[i]
                K, Ni = lv['K'], lv['N']
                # 1. post-pool GCN
                gl, dxp = E.stack_backward(ws, lv['c_post'], None if dz_next is None else dz_next.data_ptr(), Fw,
                                           dout_p + (i + 1) * Fw * 4, arg_p + (i + 1) * Fw * 4, ldo, True,
                                           d_ap[i], prec)
                put(plan.post[i], gl)
                if dxp_extra[i] is not None:
                    call('gp_axpy_f32', dxp_extra[i].data_ptr(), dxp.data_ptr(), C.c_longlong(dxp.numel()),
                         C.c_float(1.0), st)
                # 2. pooling
                dz = ws.f(B, Ni, Fw)
                if i == 0 and dS0 is not None:
                    ds, acc_ds = E._chk(dS0, 'grad of assign_tensor'), 1
                else:
                    ds, acc_ds = ws.f(B, Ni, K), 0
                wsp = ws.f(B, Ni, K)
                call('gp_pool_bwd', dxp.data_ptr(), d_ap[i].data_ptr(), lv['S'].data_ptr(), lv['z'].data_ptr(), Fw,
                     lv['adj'].data_ptr(), lv['t'].data_ptr(), E._p(lv['nb']), B, Ni, K, Fw, dz.data_ptr(), Fw, 0,
                     ds.data_ptr(), acc_ds, None if i == 0 else d_ap[i - 1].data_ptr(), wsp.data_ptr(), prec, st)
                # 3. assignment head: softmax -> Linear
                dt = ws.f(B, Ni, K)
                call('gp_softmax_mask_bwd', lv['S'].data_ptr(), ds.data_ptr(), E._p(lv['nb']), B, Ni, K,
                     dt.data_ptr(), st)
                dwp, dbp, dza = E.linear_bwd(ws, dt, lv['za'].data_ptr(), lv['Fa'], B * Ni, lv['wp'], lv['has_bp'],
                                             True)
                iw, ib = plan.assign_pred[i]
                grads[iw] = dwp
                if ib is not None:
                    grads[ib] = dbp
                # 4. assignment GCN
                gl, dxa = E.stack_backward(ws, lv['c_as'], dza.data_ptr(), lv['Fa'], None, None, 0, i > 0,
                                           None if i == 0 else d_ap[i - 1], prec)
                put(plan.assign[i], gl)
                if i > 0:
                    dxp_extra[i - 1] = dxa
                dz_next = dz
            dz_dense = dz_next
        gl, _ = E.stack_backward(ws, tape['emb'], None if dz_dense is None else dz_dense.data_ptr(), Fw, dout_p,
                                 arg_p, ldo, False, None, prec)
        put(plan.emb, gl)
        return (None, None, None, None, None) + tuple(_deliver(plan, params, grads))


class _LossFn(torch.autograd.Function):
    """CE (encoders.py:1127) [+ link loss: masked BCE (encoders.py:1311-1331) or the Frobenius option]
    [+ entropy_weight * row entropy of S (north-star option)]."""

    @staticmethod
    def forward(ctx, plan, ypred, label, S, adj):
        st = E._stream()
        ws = E.Workspace(ypred.device)
        B, Cc = ypred.shape
        ypred = E._chk(ypred, 'pred')
        if label.dtype != torch.int64 or not label.is_cuda:
            raise ValueError('label must be a CUDA int64 tensor')
        ce, probs = ws.f(1), ws.f(B, Cc)
        call('gp_ce_fwd', ypred.data_ptr(), label.data_ptr(), B, Cc, ce.data_ptr(), probs.data_ptr(), st)
        ctx.ce_scale = float(getattr(plan, 'ce_scale', 1.0))
        if ctx.ce_scale != 1.0:                      # data-parallel CE weight shard_size / global_batch (dp.py)
            call('gp_axpy_f32', ce.data_ptr(), ce.data_ptr(), C.c_longlong(1), C.c_float(ctx.ce_scale - 1.0), st)
        ctx.probs, ctx.label, ctx.B, ctx.C = probs, label, B, Cc
        link_kind = getattr(plan, 'link_kind', None) if adj is not None else None     # None | 'bce' | 'frobenius'
        ent_w = float(getattr(plan, 'ent_w', 0.0))
        ctx.link_kind, ctx.ent_w = link_kind, ent_w
        if S is None or (link_kind is None and ent_w == 0.0):
            ctx.link_kind, ctx.ent_w = None, 0.0
            return (ce.view(()),)
        # tensor-core mode with a padded cluster count: the kernels see the full [B,N,Kp] buffer (dead columns are 0),
        # the caller's S is its first K columns
        ctx.Kvis = S.shape[2]
        if getattr(plan, 'S_full', None) is not None:
            S = plan.S_full
        Bn, N, K = S.shape
        need_grad = ctx.needs_input_grad[3]
        ctx.enc_plan = getattr(plan, 'enc_plan', None)
        ctx.sb = getattr(plan, 'sb0', None)
        ctx.asym = getattr(plan, 'asym', None)
        ctx.S, ctx.nb, ctx.gsym = S, plan.nb_dev, None
        total, link, ent = ce, None, None
        if link_kind is not None:
            frob = link_kind == 'frobenius'
            total, link = ws.f(1), ws.f(1)
            ctx.pk = None
            if getattr(plan, 'packed', False) and not frob and ctx.sb is None:
                # packed schedule: the real n_b x n_b blocks only, P and dl/dP never leave shared memory
                if getattr(plan, 'inv_dev', None) is not None:
                    inv, inv_dev = 1.0, plan.inv_dev
                else:
                    inv, inv_dev = 1.0 / float(plan.num_entries), None
                total, link = PK.link_forward(ws, S, adj, plan.nb_dev, inv, inv_dev, ce)
                ctx.pk = (adj, inv, inv_dev) if need_grad else None
                npart = 0
            elif ctx.sb is not None:                    # GP_BF16: P = S S^T on tensor cores, loss in the epilogue
                partial, npart, gsym, ctx.g_upper = T.linkloss_forward(ws, ctx.sb, plan.adjb, plan.nb_dev, Bn, N, K,
                                                                       need_grad, mode=int(frob),
                                                                       adj_flags=getattr(plan, 'adj_flags', None))
            else:
                nt = (N + 63) // 64
                npart = Bn * nt * nt
                partial = ws.f(npart + 256)
                gsym = ws.f(Bn, N, N) if need_grad else None
                call('gp_frob_link_fwd' if frob else 'gp_linkloss_fwd', S.data_ptr(), adj.data_ptr(),
                     E._p(plan.nb_dev), Bn, N, K, partial.data_ptr(), E._p(gsym), st)
            if npart == 0:
                gsym = None
            elif frob:
                ctx.coef = ws.f(Bn)
                call('gp_frob_finalize', partial.data_ptr(), npart // Bn, Bn, ce.data_ptr(), total.data_ptr(),
                     link.data_ptr(), ws.f(Bn).data_ptr(), ctx.coef.data_ptr(), st)
            elif getattr(plan, 'inv_dev', None) is not None:   # 1 / sum n_b^2 lives on the device (graphed.py)
                ctx.inv, ctx.inv_dev = 1.0, plan.inv_dev
                raw = ws.f(1)
                call('gp_loss_finalize', partial.data_ptr(), npart, C.c_double(1.0), None, None, raw.data_ptr(), st)
                call('gp_mul_add_dev', raw.data_ptr(), ctx.inv_dev.data_ptr(), ce.data_ptr(), total.data_ptr(),
                     link.data_ptr(), st)
            else:
                ctx.inv, ctx.inv_dev = 1.0 / float(plan.num_entries), None
                call('gp_loss_finalize', partial.data_ptr(), npart, C.c_double(ctx.inv), ce.data_ptr(),
                     total.data_ptr(), link.data_ptr(), st)
            ctx.gsym = gsym
        if ent_w != 0.0:
            npe = int(T.load().gp_entropy_partials(Bn, N))
            pe, ent, tot2 = ws.f(npe + 256), ws.f(1), ws.f(1)
            call('gp_entropy_fwd', S.data_ptr(), E._p(plan.nb_dev), Bn, N, K, pe.data_ptr(), st)
            ctx.ent_scale = ent_w / float(plan.num_real_rows)
            call('gp_loss_finalize', pe.data_ptr(), npe, C.c_double(1.0 / float(plan.num_real_rows)), None, None,
                 ent.data_ptr(), st)
            call('gp_add_scaled', total.data_ptr(), ent.data_ptr(), C.c_float(ent_w), tot2.data_ptr(), st)
            total = tot2
        outs = [total.view(())]
        for t in (link, ent):
            if t is not None:
                t = t.view(())
                ctx.mark_non_differentiable(t)
                outs.append(t)
        ctx.n_out = len(outs)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, g, *_unused):
        st = E._stream()
        ws = E.Workspace(g.device)
        g = E._chk(g, 'grad of loss')
        dy = ws.f(ctx.B, ctx.C)
        call('gp_ce_bwd', ctx.probs.data_ptr(), ctx.label.data_ptr(), g.data_ptr(), ctx.B, ctx.C, dy.data_ptr(), st)
        if ctx.ce_scale != 1.0:
            call('gp_axpy_f32', dy.data_ptr(), dy.data_ptr(), C.c_longlong(dy.numel()), C.c_float(ctx.ce_scale - 1.0), st)
        dS = None
        if ctx.link_kind is not None and getattr(ctx, 'pk', None) is not None:
            adj_, inv_, inv_dev_ = ctx.pk
            dS = PK.link_backward(ws, ctx.S, adj_, ctx.nb, inv_, inv_dev_, g)
            ctx.pk = None
        if ctx.link_kind is not None and ctx.gsym is not None:
            S = ctx.S
            Bn, N, K = S.shape
            frob = ctx.link_kind == 'frobenius'
            lim = ctx.nb is not None
            if not frob and getattr(ctx, 'inv_dev', None) is not None:       # upstream * (1 / sum n_b^2), on device
                gi = ws.f(1)
                call('gp_mul_add_dev', g.data_ptr(), ctx.inv_dev.data_ptr(), None, None, gi.data_ptr(), st)
                g_link = gi
            else:
                g_link = g
            if ctx.sb is not None:
                if frob:        # per-graph factor coef[b] * upstream folded into a scaled bf16 copy of S
                    ssb = T.bfbuf(ws, Bn, N, K)
                    call('gp_scale_rows_batch', S.data_ptr(), ctx.coef.data_ptr(), g.data_ptr(), Bn, N, K, None, 0,
                         ssb.ptr, ssb.ld, ssb.ld, st)
                    dS = T.linkloss_backward(ws, ctx.gsym, ssb, ctx.nb, Bn, N, K, 1.0, None, asym=ctx.asym)
                else:
                    dS = T.linkloss_backward(ws, ctx.gsym, ctx.sb, ctx.nb, Bn, N, K, ctx.inv, g_link.data_ptr(), asym=ctx.asym,
                                             upper=getattr(ctx, 'g_upper', False))
            else:
                dS = ws.f(Bn, N, K)
                if frob:
                    ss = ws.f(Bn, N, K)
                    call('gp_scale_rows_batch', S.data_ptr(), ctx.coef.data_ptr(), g.data_ptr(), Bn, N, K,
                         ss.data_ptr(), K, None, 0, K, st)
                    E.bgemm(ctx.gsym.data_ptr(), ss.data_ptr(), dS.data_ptr(), N, K, N, Bn, (N * N, N, 1),
                            (N * K, K, 1), (N * K, K, 1), lim=E._p(ctx.nb), lim_m=int(lim), lim_k=int(lim))
                else:
                    E.bgemm(ctx.gsym.data_ptr(), S.data_ptr(), dS.data_ptr(), N, K, N, Bn, (N * N, N, 1),
                            (N * K, K, 1), (N * K, K, 1), lim=E._p(ctx.nb), lim_m=int(lim), lim_k=int(lim),
                            alpha=ctx.inv, alpha_dev=g_link.data_ptr())
            ctx.gsym = None
        if ctx.ent_w != 0.0 and ctx.needs_input_grad[3]:
            S = ctx.S
            Bn, N, K = S.shape
            acc = dS is not None
            if dS is None:
                dS = ws.f(Bn, N, K)
            call('gp_entropy_bwd', S.data_ptr(), E._p(ctx.nb), Bn, N, K, g.data_ptr(), C.c_float(ctx.ent_scale),
                 dS.data_ptr(), int(acc), st)
        if dS is not None and dS.shape[2] != ctx.Kvis:
            if ctx.enc_plan is not None:
                ctx.enc_plan.ds_tag = dS.data_ptr()      # _padded_grad recognises the buffer and uses it in place
            dS = dS[:, :, :ctx.Kvis]
        return None, dy, None, dS, None


class _LinkHopFn(torch.autograd.Function):
    """Link-prediction loss with adj_hop > 1 (encoders.py:1312-1317): the predicted adjacency is
    Q = sum_{h=1..hop} P^h with P = S S^T, clamped at 1 (R3).  Never passed by the reference's callers, so it runs on
    the fp32 FFMA schedule in either precision mode: the powers of P are materialised ([B,N,N] fp32 each, like the
    reference does) by the library's batched GEMM, gp_linkloss_from_q does clamp + masked BCE + G = dl/dQ, and the
    backward pushes G through the powers:  D_1 = G,  D_h = P D_{h-1} + G P^{h-1},  dP = sum_h D_h,
    dS = (dP + dP^T) S.  forward returns (base + link, link); `base` is the CE (+ entropy) scalar."""

    @staticmethod
    def forward(ctx, S, base, adj, nb_dev, inv, hop):
        st = E._stream()
        ws = E.Workspace(S.device)
        S = E._chk(S, 'assign_tensor')
        B, N, K = S.shape
        nbp, lim = E._p(nb_dev), int(nb_dev is not None)
        sNN, sNK = (N * N, N, 1), (N * K, K, 1)
        P1 = ws.f(B, N, N)
        E.bgemm(S.data_ptr(), S.data_ptr(), P1.data_ptr(), N, N, K, B, sNK, (N * K, 1, K), sNN, lim=nbp, lim_m=lim,
                lim_n=lim)
        pows = [P1]
        Q = ws.f(B, N, N)
        call('gp_pad_copy_f32', P1.data_ptr(), C.c_longlong(N), C.c_longlong(B * N), N, Q.data_ptr(),
             C.c_longlong(N), C.c_longlong(B * N), N, C.c_float(0.0), st)
        for _ in range(hop - 1):
            Th = ws.f(B, N, N)
            E.bgemm(pows[-1].data_ptr(), P1.data_ptr(), Th.data_ptr(), N, N, N, B, sNN, sNN, sNN, lim=nbp, lim_m=lim,
                    lim_n=lim, lim_k=lim)
            call('gp_axpy_f32', Th.data_ptr(), Q.data_ptr(), C.c_longlong(Q.numel()), C.c_float(1.0), st)
            pows.append(Th)
        need_grad = ctx.needs_input_grad[0]
        npart = int(T.load().gp_linkloss_from_q_partials(B, N))
        partial = ws.f(npart + 256)
        G = ws.f(B, N, N) if need_grad else None
        call('gp_linkloss_from_q', Q.data_ptr(), adj.data_ptr(), nbp, B, N, partial.data_ptr(), E._p(G), st)
        total, link = ws.f(1), ws.f(1)
        call('gp_loss_finalize', partial.data_ptr(), npart, C.c_double(inv), base.data_ptr(), total.data_ptr(),
             link.data_ptr(), st)
        ctx.tape = (S, pows[:-1], G, nb_dev, float(inv), int(hop))
        link = link.view(())
        ctx.mark_non_differentiable(link)
        return total.view(()), link

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _unused=None):
        S, pows, G, nb_dev, inv, hop = ctx.tape
        ctx.tape = None
        st = E._stream()
        ws = E.Workspace(S.device)
        g = E._chk(g, 'grad of loss')
        dS = None
        if G is not None:
            B, N, K = S.shape
            nbp, lim = E._p(nb_dev), int(nb_dev is not None)
            sNN, sNK = (N * N, N, 1), (N * K, K, 1)
            kw = dict(lim=nbp, lim_m=lim, lim_n=lim, lim_k=lim)
            D, tot = G, G
            if hop > 1:
                tot = ws.f(B, N, N)
                call('gp_pad_copy_f32', G.data_ptr(), C.c_longlong(N), C.c_longlong(B * N), N, tot.data_ptr(),
                     C.c_longlong(N), C.c_longlong(B * N), N, C.c_float(0.0), st)
            for h in range(2, hop + 1):
                Dn = ws.f(B, N, N)
                E.bgemm(pows[0].data_ptr(), D.data_ptr(), Dn.data_ptr(), N, N, N, B, sNN, sNN, sNN, **kw)
                E.bgemm(G.data_ptr(), pows[h - 2].data_ptr(), Dn.data_ptr(), N, N, N, B, sNN, sNN, sNN, beta=1.0, **kw)
                call('gp_axpy_f32', Dn.data_ptr(), tot.data_ptr(), C.c_longlong(tot.numel()), C.c_float(1.0), st)
                D = Dn
            dS = ws.f(B, N, K)
            E.bgemm(tot.data_ptr(), S.data_ptr(), dS.data_ptr(), N, K, N, B, sNN, sNK, sNK, lim=nbp, lim_m=lim,
                    lim_k=lim, alpha=inv, alpha_dev=g.data_ptr())
            E.bgemm(tot.data_ptr(), S.data_ptr(), dS.data_ptr(), N, K, N, B, (N * N, 1, N), sNK, sNK, lim=nbp,
                    lim_m=lim, lim_k=lim, alpha=inv, alpha_dev=g.data_ptr(), beta=1.0)
        return dS, g, None, None, None, None


# ------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------
class GraphConv(nn.Module):
    """R1: the DiffPool GraphConv of encoders.py:296-328 (weight is [in, out], x@W layout)."""

    def __init__(self, input_dim, output_dim, add_self=False, normalize_embedding=False,
                 dropout=0.0, bias=True):
        super().__init__()
        self.add_self = add_self
        self.dropout = dropout
        if dropout > 0.001:
            self.dropout_layer = nn.Dropout(p=dropout)
        self.normalize_embedding = normalize_embedding
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.weight = nn.Parameter(torch.empty(input_dim, output_dim))
        nn.init.xavier_uniform_(self.weight.data, gain=nn.init.calculate_gain('relu'))
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_dim))
        else:
            self.bias = None

    def forward(self, x, adj):
        if self.dropout > 0.001 and self.training:                            # encoders.py:316-317
            x = _DropoutFn.apply(E._chk(x, 'x'), float(self.dropout), _draw_seed())
        return _GraphConvFn.apply(x, adj, self.weight, self.bias, self.add_self, self.normalize_embedding)


def _draw_seed():
    """Base seed of one forward call's dropout masks, from torch's CPU generator (torch.manual_seed reproduces it)."""
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        raise NotImplementedError('gp_b200: dropout inside a captured CUDA graph would replay one fixed mask')
    return int(torch.randint(0, 1 << 62, (1,)).item())


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        y = torch.empty_like(x)
        d = x.shape[-1]
        E.dropout(x.data_ptr(), d, x.numel() // d, d, p, seed, y.data_ptr(), d)
        ctx.ps = (p, seed)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        dy = E._chk(dy, 'dy')
        dx = torch.empty_like(dy)
        d = dy.shape[-1]
        E.dropout(dy.data_ptr(), d, dy.numel() // d, d, ctx.ps[0], ctx.ps[1], dx.data_ptr(), d)
        return dx, None, None


class _GraphConvFn(torch.autograd.Function):
    """Stand-alone GraphConv (the encoders below run whole stacks through _EncoderFn instead)."""

    @staticmethod
    def forward(ctx, x, adj, w, b, add_self, normalize):
        x, adj = E._chk(x, 'x'), E._chk(adj, 'adj')
        B, N, din = x.shape
        dout = w.shape[1]
        ws = E.Workspace(x.device)
        u, y, rn = ws.f(B, N, din), ws.f(B, N, dout), ws.f(B, N)
        call('gp_graphconv_fwd', x.data_ptr(), din, adj.data_ptr(), w.data_ptr(), E._p(b), None, B, N, din, dout,
             int(add_self), int(normalize), u.data_ptr(), y.data_ptr(), dout, rn.data_ptr(), E.F32, E._stream())
        # y is this node's output: keep a detached alias (saving the output object itself would form a cycle)
        ctx.saved = (x, adj, w, b is not None, u, y.detach(), rn, add_self, normalize)
        ctx.need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, adj, w, has_b, u, y, rn, add_self, normalize = ctx.saved
        B, N, din = x.shape
        dout = w.shape[1]
        ws = E.Workspace(x.device)
        dy = E._chk(dy, 'dy')
        dv = ws.f(B, N, dout)
        call('gp_gcn_layer_bwd', dy.data_ptr(), dout, None, None, None, 0, None, 0, y.data_ptr(), dout,
             rn.data_ptr(), None, B, N, dout, 0, 0, int(normalize), dv.data_ptr(), E._stream())
        need_dx, need_da = ctx.need
        dw, db = ws.f(din, dout), (ws.f(dout) if has_b else None)
        du = ws.f(B, N, din) if (need_dx or need_da) else None
        dx = ws.f(B, N, din) if need_dx else None
        da = ws.z(B, N, N) if need_da else None
        cs = ws.f(int(T.load().gp_graphconv_bwd_ws(B, N, din, dout, int(add_self))))
        call('gp_graphconv_bwd', dv.data_ptr(), u.data_ptr(), x.data_ptr(), din, adj.data_ptr(), w.data_ptr(), None,
             B, N, din, dout, int(add_self), dw.data_ptr(), E._p(db), E._p(du), E._p(dx), E._p(da), cs.data_ptr(),
             E.F32, E._stream())
        return dx, da, dw, db, None, None


class GcnEncoderGraph(nn.Module):
    """encoders.py:976-1134 (method=base)."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 pred_hidden_dims=[], concat=True, bn=True, dropout=0.0, args=None):
        super().__init__()
        self.concat = concat
        add_self = not concat
        self.bn = bn
        self.num_layers = num_layers
        self.num_aggs = 1
        self.precision = DEFAULT_PRECISION
        self._ce_scale, self._entries_override = 1.0, None
        self.bias = True
        if args is not None:
            self.bias = args.bias
        self.conv_first, self.conv_block, self.conv_last = self.build_conv_layers(
            input_dim, hidden_dim, embedding_dim, num_layers, add_self, normalize=True, dropout=dropout)
        self.act = nn.ReLU()
        self.label_dim = label_dim
        if concat:
            self.pred_input_dim = hidden_dim * (num_layers - 1) + embedding_dim
        else:
            self.pred_input_dim = embedding_dim
        self.pred_model = self.build_pred_layers(self.pred_input_dim, pred_hidden_dims, label_dim,
                                                 num_aggs=self.num_aggs)
        self._reinit()

    def _reinit(self):
        for m in self.modules():                                             # encoders.py:1003-1007
            if isinstance(m, GraphConv):
                nn.init.xavier_uniform_(m.weight.data, gain=nn.init.calculate_gain('relu'))
                if m.bias is not None:
                    nn.init.constant_(m.bias.data, 0.0)

    def build_conv_layers(self, input_dim, hidden_dim, embedding_dim, num_layers, add_self,
                          normalize=False, dropout=0.0):
        conv_first = GraphConv(input_dim=input_dim, output_dim=hidden_dim, add_self=add_self,
                               normalize_embedding=normalize, bias=self.bias)
        conv_block = nn.ModuleList(
            [GraphConv(input_dim=hidden_dim, output_dim=hidden_dim, add_self=add_self,
                       normalize_embedding=normalize, dropout=dropout, bias=self.bias)
             for _ in range(num_layers - 2)])
        conv_last = GraphConv(input_dim=hidden_dim, output_dim=embedding_dim, add_self=add_self,
                              normalize_embedding=normalize, bias=self.bias)
        return conv_first, conv_block, conv_last

    def build_pred_layers(self, pred_input_dim, pred_hidden_dims, label_dim, num_aggs=1):
        pred_input_dim = pred_input_dim * num_aggs
        if len(pred_hidden_dims) == 0:
            return nn.Linear(pred_input_dim, label_dim)
        layers = []
        for pred_dim in pred_hidden_dims:
            layers.append(nn.Linear(pred_input_dim, pred_dim))
            layers.append(self.act)
            pred_input_dim = pred_dim
        layers.append(nn.Linear(pred_dim, label_dim))
        return nn.Sequential(*layers)

    def construct_mask(self, max_nodes, batch_num_nodes):
        """Kept for API compatibility (encoders.py:1035-1046); the kernels never materialise it."""
        nb = torch.as_tensor(np.asarray(batch_num_nodes).astype(np.int64))
        return (torch.arange(max_nodes)[None, :] < nb[:, None]).float().unsqueeze(2)

    # ---- plan construction -----------------------------------------------------------------
    def _conv_pairs(self, params, first, block, last):
        pairs = []
        for m in [first] + list(block) + [last]:
            params.append(m.weight)
            iw = len(params) - 1
            ib = None
            if m.bias is not None:
                params.append(m.bias)
                ib = len(params) - 1
            pairs.append((iw, ib))
        return pairs

    def _drops(self, block):
        """Dropout probability per layer of a stack: only conv_block layers have one (encoders.py:1015), and only in
        training mode (nn.Dropout)."""
        mid = [float(m.dropout) if (self.training and m.dropout > 0.001) else 0.0 for m in block]
        return [0.0] + mid + [0.0]

    def _pred_pairs(self, params, model):
        mods = [model] if isinstance(model, nn.Linear) else [m for m in model if isinstance(m, nn.Linear)]
        pairs = []
        for m in mods:
            params.append(m.weight)
            iw = len(params) - 1
            ib = None
            if m.bias is not None:
                params.append(m.bias)
                ib = len(params) - 1
            pairs.append((iw, ib))
        return pairs

    def _base_plan(self, x, adj, batch_num_nodes):
        if torch.is_tensor(x) and x.dtype == torch.bfloat16:
            # compact feed (tensor-core mode only): bf16 features are what the contractions read anyway; the row must
            # satisfy TMA's 16-byte rule
            if not (self.precision == T.BF16 and self.concat and x.is_cuda and x.is_contiguous() and x.dim() == 3
                    and x.shape[2] % 8 == 0):
                raise ValueError('bfloat16 features need the tensor-core mode, a contiguous CUDA tensor and D % 8 == 0')
        else:
            x = E._chk(x, 'x')
        if isinstance(adj, T.PreparedAdjacency):             # bf16 operand built by the feed (tensor-core mode only)
            if not (self.precision == T.BF16 and self.concat):
                raise ValueError('a PreparedAdjacency needs the tensor-core mode (model.precision = 1)')
        else:
            adj = E._chk_adj(adj, self.precision == T.BF16 and self.concat)
        if len(adj.shape) != 3:
            raise ValueError('adj must be [B,N,N]')
        if x.dim() != 3 or adj.shape[1] != adj.shape[2] or tuple(adj.shape[:2]) != tuple(x.shape[:2]):
            raise ValueError('expected x [B,N,D] and adj [B,N,N], got %s and %s' % (tuple(x.shape), tuple(adj.shape)))
        if x.shape[2] != self.conv_first.weight.shape[0]:
            raise ValueError('input feature dim %d != conv_first input dim %d'
                             % (x.shape[2], self.conv_first.weight.shape[0]))
        plan = _Plan()
        plan.precision = self.precision
        plan.grad_sink = getattr(self, '_grad_sink', None)
        plan.nb_dev, plan.nb_host = E.prep_nb(batch_num_nodes, adj.shape[1], x.device)
        if plan.nb_host is not None and len(plan.nb_host) != x.shape[0]:
            raise ValueError('batch_num_nodes has %d entries for a batch of %d' % (len(plan.nb_host), x.shape[0]))
        plan.concat = self.concat
        plan.add_self = not self.concat
        if plan.add_self:
            plan.precision = E.F32           # the tensor-core schedule covers add_self=False (concat) only
        plan.bn = self.bn
        plan.label_dim = self.label_dim
        params = []
        plan.emb = self._conv_pairs(params, self.conv_first, self.conv_block, self.conv_last)
        plan.emb_drop = self._drops(self.conv_block)
        plan.post_drop = []
        plan.seed = _draw_seed() if any(plan.emb_drop) else 0
        plan.F = sum(params[iw].shape[1] for iw, _ in plan.emb)
        plan.douts_last = self.conv_last.weight.shape[1]
        return plan, params, x, adj

    def forward(self, x, adj, batch_num_nodes=None, **kwargs):
        plan, params, x, adj = self._base_plan(x, adj, batch_num_nodes)
        plan.soft = False
        plan.num_pooling = 0
        plan.pred = self._pred_pairs(params, self.pred_model)
        self._plan = plan
        return _apply_encoder(plan, x, adj, None, params)

    def loss(self, pred, label, type='softmax'):
        if type != 'softmax':
            raise NotImplementedError("gp_b200: only type='softmax' is implemented (callers never pass 'margin')")
        plan = _Plan()
        plan.ce_scale = self._ce_scale
        return _LossFn.apply(plan, pred, label, None, None)[0]

    def set_loss_scaling(self, ce_scale=1.0, num_entries=None):
        """Data-parallel hook (dp.py, mode='global_norm'): weight of this shard's CE mean and the GLOBAL sum of
        n_b^2 that normalises the link loss instead of the shard-local one (encoders.py:1326)."""
        self._ce_scale = float(ce_scale)
        self._entries_override = None if num_entries is None else int(num_entries)


class Set2Set(nn.Module):
    """set2set.py:8-57.  Holds the reference's parameters under the reference's names (``lstm.weight_ih_l0`` ...,
    ``pred.weight``); the nn.LSTM / nn.Linear objects are parameter containers only -- the n sequential LSTM +
    attention steps run in gp_set2set_fwd / gp_set2set_bwd (one CTA per graph), the projection on the library's GEMM."""

    def __init__(self, input_dim, hidden_dim, act_fn=nn.ReLU, num_layers=1):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError('gp_b200: Set2Set with num_layers != 1 (the reference always uses 1)')
        if act_fn is not nn.ReLU:
            raise NotImplementedError('gp_b200: Set2Set supports act_fn=nn.ReLU (the reference default)')
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        if hidden_dim <= input_dim:
            print('ERROR: Set2Set output_dim should be larger than input_dim')
        self.lstm_output_dim = hidden_dim - input_dim
        if hidden_dim != 2 * input_dim:
            # q* = [q, r] has lstm_output_dim + input_dim entries and e = E q^T needs lstm_output_dim == input_dim
            raise ValueError('Set2Set needs hidden_dim == 2 * input_dim (set2set.py:51,55)')
        self.lstm = nn.LSTM(hidden_dim, input_dim, num_layers=num_layers, batch_first=True)
        self.pred = nn.Linear(hidden_dim, input_dim)
        self.act = act_fn()

    def _params(self):
        l = self.lstm
        return [l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0]

    def forward(self, embedding, batch_num_nodes=None):
        """[B,n,d] -> [B,d].  batch_num_nodes (an extension): rows beyond a graph's node count are treated as zero rows."""
        emb = E._chk(embedding, 'embedding')
        nb_dev, _ = E.prep_nb(batch_num_nodes, emb.shape[1], emb.device)
        return _Set2SetFn.apply(emb, nb_dev, self.pred.weight, self.pred.bias, *self._params())


def _s2s_forward(ws, e_ptr, lde, nb, B, N, d, lstm_params):
    wih, whh, bih, bhh = lstm_params
    qs, gates = ws.f(B, N + 1, 2 * d), ws.f(B, N, 4 * d)
    cells, att = ws.f(B, N, d), ws.f(B, N, N)
    call('gp_set2set_fwd', e_ptr, C.c_longlong(lde), E._p(nb), B, N, d, wih.data_ptr(), whh.data_ptr(), E._p(bih),
         E._p(bhh), qs.data_ptr(), gates.data_ptr(), cells.data_ptr(), att.data_ptr(), E._stream())
    return qs, gates, cells, att


def _s2s_backward(ws, e_ptr, lde, nb, B, N, d, lstm_params, saved, dqs):
    """BPTT kernel + the batched contractions that turn its per-step gradients into parameter / embedding gradients.
    Returns (dW_ih, dW_hh, db_ih, db_hh, dE [B,N,d])."""
    st = E._stream()
    wih, whh, bih, bhh = lstm_params
    qs, gates, cells, att = saved
    dz, dr, de = ws.f(B, N + 1, 4 * d), ws.f(B, N, d), ws.f(B, N, N)
    call('gp_set2set_bwd', e_ptr, C.c_longlong(lde), E._p(nb), B, N, d, wih.data_ptr(), whh.data_ptr(),
         gates.data_ptr(), cells.data_ptr(), att.data_ptr(), dqs.data_ptr(), C.c_longlong(2 * d), dz.data_ptr(),
         dr.data_ptr(), de.data_ptr(), st)
    rows = B * (N + 1)
    split = max(1, min(512, rows // 1024))     # split-K over the rows: enough CTAs to stream them at HBM rate
    dwih, dwhh = ws.f(4 * d, 2 * d), ws.f(4 * d, d)
    # dW_ih = DZ^T QS ;  dW_hh = DZ^T QS[:, :d]  (h_{t-1} is the first half of q*_{t-1})
    E.bgemm(dz.data_ptr(), qs.data_ptr(), dwih.data_ptr(), 4 * d, 2 * d, rows, 1, (0, 1, 4 * d), (0, 2 * d, 1),
            (0, 2 * d, 1), split_k=split)
    E.bgemm(dz.data_ptr(), qs.data_ptr(), dwhh.data_ptr(), 4 * d, d, rows, 1, (0, 1, 4 * d), (0, 2 * d, 1),
            (0, d, 1), split_k=split)
    dbs = []
    for b_ in (bih, bhh):                     # the two bias vectors enter the gates as a sum: same gradient, own buffer
        if b_ is None:
            dbs.append(None)
            continue
        db, cs = ws.f(4 * d), ws.f(256 * 4 * d)
        call('gp_colsum_f32', dz.data_ptr(), C.c_longlong(rows), 4 * d, C.c_longlong(4 * d), db.data_ptr(), 0,
             cs.data_ptr(), st)
        dbs.append(db)
    # dE_b = AT_b^T DR_b + DE_b^T Q_b, rows beyond n_b stay zero (the mask of encoders.py:1080)
    dE = ws.f(B, N, d)
    lim = int(nb is not None)
    E.bgemm(att.data_ptr(), dr.data_ptr(), dE.data_ptr(), N, d, N, B, (N * N, 1, N), (N * d, d, 1), (N * d, d, 1),
            lim=E._p(nb), lim_m=lim)
    E.bgemm(de.data_ptr(), qs.data_ptr() + 2 * d * 4, dE.data_ptr(), N, d, N, B, (N * N, 1, N),
            ((N + 1) * 2 * d, 2 * d, 1), (N * d, d, 1), lim=E._p(nb), lim_m=lim, beta=1.0)
    return dwih, dwhh, dbs[0], dbs[1], dE


class _Set2SetFn(torch.autograd.Function):
    """Stand-alone Set2Set module: out = relu(pred(q*_n))."""

    @staticmethod
    def forward(ctx, emb, nb, wp, bp, wih, whh, bih, bhh):
        ws = E.Workspace(emb.device)
        B, N, d = emb.shape
        saved = _s2s_forward(ws, emb.data_ptr(), d, nb, B, N, d, (wih, whh, bih, bhh))
        qs = saved[0]
        out = E.linear_fwd(ws, qs.data_ptr() + N * 2 * d * 4, (N + 1) * 2 * d, B, wp, bp, relu=True)
        ctx.tape = (emb, nb, wp, bp is not None, (wih, whh, bih, bhh), saved, out.detach())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        emb, nb, wp, has_bp, lstm_params, saved, out = ctx.tape
        ctx.tape = None
        ws = E.Workspace(emb.device)
        B, N, d = emb.shape
        dout = E._chk(dout, 'grad of the Set2Set output')
        g = ws.f(B, d)
        call('gp_relu_mask_bwd', dout.data_ptr(), out.data_ptr(), C.c_longlong(g.numel()), g.data_ptr(), E._stream())
        qs = saved[0]
        dwp, dbp, dqs = E.linear_bwd(ws, g, qs.data_ptr() + N * 2 * d * 4, (N + 1) * 2 * d, B, wp, has_bp, True)
        dwih, dwhh, dbih, dbhh, dE = _s2s_backward(ws, emb.data_ptr(), d, nb, B, N, d, lstm_params, saved, dqs)
        return dE, None, dwp, dbp, dwih, dwhh, dbih, dbhh


class _S2SEncoderFn(torch.autograd.Function):
    """GcnSet2SetEncoder.forward (encoders.py:1144-1157) as one autograd node: GCN stack (fp32 schedule) -> masked
    concat -> Set2Set -> relu(s2s.pred) -> pred_model."""

    @staticmethod
    def forward(ctx, plan, x, adj, *params):
        ws = E.Workspace(x.device)
        B, N, D = x.shape
        nb = plan.nb_dev
        w0 = [_wb(params, p)[0] for p in plan.emb]
        b0 = [_wb(params, p)[1] for p in plan.emb]
        z, c_emb = E.stack_forward(ws, x.data_ptr(), D, D, adj, nb, B, N, w0, b0, plan.add_self, plan.bn, E.F32,
                                   drops=plan.emb_drop, seed=plan.seed)
        d = plan.F
        lstm_params = tuple(None if i is None else params[i] for i in plan.lstm)
        saved = _s2s_forward(ws, z.data_ptr(), d, nb, B, N, d, lstm_params)
        lin = [_wb(params, p) for p in plan.pred]         # [s2s.pred (ReLU after it), pred_model ...]
        ypred, acts = E.mlp_fwd(ws, saved[0].data_ptr() + N * 2 * d * 4, (N + 1) * 2 * d, B, lin)
        ctx.tape = dict(plan=plan, params=params, B=B, N=N, emb=c_emb, z=z, saved=saved, acts=acts, lin=lin,
                        lstm=lstm_params)
        return ypred

    @staticmethod
    @once_differentiable
    def backward(ctx, dypred):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError('gp_b200: backward called twice (buffers were freed)')
        ctx.tape = None
        plan, params, B, N = tape['plan'], tape['params'], tape['B'], tape['N']
        ws = E.Workspace(tape['z'].device)
        d = plan.F
        grads = [None] * len(params)
        dypred = E._chk(dypred, 'grad of ypred')
        dqs = ws.f(B, 2 * d)
        gl = E.mlp_bwd(ws, dypred, B, tape['acts'], tape['lin'], dqs.data_ptr(), 2 * d)
        for (iw, ib), (dw, db) in zip(plan.pred, gl):
            grads[iw] = dw
            if ib is not None:
                grads[ib] = db
        dwih, dwhh, dbih, dbhh, dE = _s2s_backward(ws, tape['z'].data_ptr(), d, plan.nb_dev, B, N, d, tape['lstm'],
                                                   tape['saved'], dqs)
        for i, gq in zip(plan.lstm, (dwih, dwhh, dbih, dbhh)):
            if i is not None:
                grads[i] = gq
        gl, _ = E.stack_backward(ws, tape['emb'], dE.data_ptr(), d, None, None, 0, False, None, E.F32)
        for (iw, ib), (dw, db) in zip(plan.emb, gl):
            grads[iw] = dw
            if ib is not None:
                grads[ib] = db
        return (None, None, None) + tuple(grads)


class GcnSet2SetEncoder(GcnEncoderGraph):
    """encoders.py:1137-1157 (method=base-set2set): GCN stack, masked concat, Set2Set readout (set2set.py), pred_model.
    Outside the north-star path: runs on the fp32 schedule whatever `precision` says."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 pred_hidden_dims=[], concat=True, bn=True, dropout=0.0, args=None):
        super().__init__(input_dim, hidden_dim, embedding_dim, label_dim, num_layers, pred_hidden_dims, concat,
                         bn, dropout, args=args)
        if not concat:
            # gcn_forward always concatenates (encoders.py:1078) while Set2Set is sized for pred_input_dim =
            # embedding_dim when concat=False (:1141): the reference cannot run that combination either
            raise NotImplementedError('gp_b200: GcnSet2SetEncoder requires concat=True (as the reference does)')
        self.s2s = Set2Set(self.pred_input_dim, self.pred_input_dim * 2)

    def forward(self, x, adj, batch_num_nodes=None, **kwargs):
        if isinstance(adj, T.PreparedAdjacency):
            raise ValueError('GcnSet2SetEncoder runs on the fp32 schedule: pass the fp32 adjacency')
        prec, self.precision = self.precision, E.F32
        try:
            plan, params, x, adj = self._base_plan(x, adj, batch_num_nodes)
        finally:
            self.precision = prec
        plan.precision = E.F32
        plan.soft, plan.num_pooling = False, 0
        plan.lstm = []
        for p in self.s2s._params():
            params.append(p)
            plan.lstm.append(len(params) - 1)
        plan.pred = self._pred_pairs(params, self.s2s.pred) + self._pred_pairs(params, self.pred_model)
        self._plan = plan
        return _S2SEncoderFn.apply(plan, x, adj, *params)


class SoftPoolingGcnEncoder(GcnEncoderGraph):
    """encoders.py:1160-1334 (method=soft-assign, DiffPool)."""

    def __init__(self, max_num_nodes, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 assign_hidden_dim, assign_ratio=0.25, assign_num_layers=-1, num_pooling=1,
                 pred_hidden_dims=[50], concat=True, bn=True, dropout=0.0, linkpred=True,
                 assign_input_dim=-1, args=None, link_loss='bce', entropy_weight=0.0):
        # link_loss / entropy_weight: north-star options appended AFTER the reference's arguments (defaults
        # reproduce the reference: masked-BCE link loss, no entropy term -- SURVEY.md appendix A.6).
        # R8: bn / dropout are NOT forwarded to the first GCN (encoders.py:1172-1173)
        super().__init__(input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                         pred_hidden_dims=pred_hidden_dims, concat=concat, args=args)
        if not concat:
            # the reference's gcn_forward always concatenates (encoders.py:1078) while the post-pool
            # GCN is sized for embedding_dim only (:1185-1186): concat=False cannot run there either.
            raise NotImplementedError('gp_b200: SoftPoolingGcnEncoder requires concat=True (as the reference does)')
        add_self = not concat
        self.num_pooling = num_pooling
        self.linkpred = linkpred
        self.assign_ent = True
        if link_loss not in ('bce', 'frobenius'):
            raise ValueError("link_loss must be 'bce' (the reference) or 'frobenius'")
        self.link_loss_kind = link_loss
        self.entropy_weight = float(entropy_weight)

        def reg(name, i, mod):                                               # R5
            setattr(self, name if i == num_pooling - 1 else '%s_l%d' % (name, i), mod)
            return mod

        self.conv_first_after_pool, self.conv_block_after_pool, self.conv_last_after_pool = [], [], []
        for i in range(num_pooling):
            f, b, l = self.build_conv_layers(self.pred_input_dim, hidden_dim, embedding_dim, num_layers,
                                             add_self, normalize=True, dropout=dropout)
            self.conv_first_after_pool.append(reg('conv_first2', i, f))
            self.conv_block_after_pool.append(reg('conv_block2', i, b))
            self.conv_last_after_pool.append(reg('conv_last2', i, l))

        if assign_num_layers == -1:
            assign_num_layers = num_layers
        if assign_input_dim == -1:
            assign_input_dim = input_dim
        self.assign_conv_first_modules, self.assign_conv_block_modules = [], []
        self.assign_conv_last_modules, self.assign_pred_modules = [], []
        self.assign_dims = []
        assign_dim = int(max_num_nodes * assign_ratio)
        for i in range(num_pooling):
            if assign_dim < 1:
                raise ValueError('assign_dim became 0 at pooling level %d' % i)
            self.assign_dims.append(assign_dim)
            f, b, l = self.build_conv_layers(assign_input_dim, assign_hidden_dim, assign_dim,
                                             assign_num_layers, add_self, normalize=True)
            apin = assign_hidden_dim * (num_layers - 1) + assign_dim if concat else assign_dim
            ap = self.build_pred_layers(apin, [], assign_dim, num_aggs=1)
            assign_input_dim = self.pred_input_dim                           # R6
            assign_dim = int(assign_dim * assign_ratio)
            self.assign_conv_first_modules.append(reg('assign_conv_first', i, f))
            self.assign_conv_block_modules.append(reg('assign_conv_block', i, b))
            self.assign_conv_last_modules.append(reg('assign_conv_last', i, l))
            self.assign_pred_modules.append(reg('assign_pred', i, ap))

        self.pred_model = self.build_pred_layers(self.pred_input_dim * (num_pooling + 1), pred_hidden_dims,
                                                 label_dim, num_aggs=self.num_aggs)
        self._reinit()

    def forward(self, x, adj, batch_num_nodes, **kwargs):
        ax_in = kwargs.get('assign_x')
        same_ax = ax_in is None or ax_in is x
        plan, params, x, adj = self._base_plan(x, adj, batch_num_nodes)
        x_a = x if same_ax else E._chk(ax_in, 'assign_x')
        if x_a.shape[:2] != x.shape[:2]:
            raise ValueError('assign_x batch/node dims differ from x')
        plan.soft = True
        plan.bn = True                   # R8: the first GCN always uses BN
        plan.bn_post = True              # post-pool stacks share self.bn == True (ctor default, :1172)
        plan.num_pooling = self.num_pooling
        plan.assign_dims = list(self.assign_dims)
        plan.post, plan.assign, plan.assign_pred = [], [], []
        for i in range(self.num_pooling):
            plan.post.append(self._conv_pairs(params, self.conv_first_after_pool[i], self.conv_block_after_pool[i],
                                              self.conv_last_after_pool[i]))
            plan.post_drop.append(self._drops(self.conv_block_after_pool[i]))
            plan.assign.append(self._conv_pairs(params, self.assign_conv_first_modules[i],
                                                self.assign_conv_block_modules[i],
                                                self.assign_conv_last_modules[i]))
            plan.assign_pred.append(self._pred_pairs(params, self.assign_pred_modules[i])[0])
        plan.pred = self._pred_pairs(params, self.pred_model)
        if any(any(d) for d in plan.post_drop) and not plan.seed:
            plan.seed = _draw_seed()
        plan.packed = PK.supported(plan, x, adj, x_a, params)
        self._plan = plan
        ypred, S0 = _apply_encoder(plan, x, adj, x_a, params)
        self.assign_tensors = [S0] + plan.all_S[1:]
        self.assign_tensor = plan.all_S[-1] if self.num_pooling > 1 else S0   # last level's S (:1269,1273)
        self._S0 = S0
        return ypred

    def loss(self, pred, label, adj=None, batch_num_nodes=None, adj_hop=1):
        plan = self._plan
        ent_w = self.entropy_weight
        if not self.linkpred and ent_w == 0.0:
            lp0 = _Plan()
            lp0.ce_scale = self._ce_scale
            return _LossFn.apply(lp0, pred, label, None, None)[0]
        adj_hop = int(adj_hop)
        if adj_hop < 1:
            raise ValueError('adj_hop must be >= 1')
        S0 = self._S0                                                        # R7: level-0 S with level-0 adj
        lp = _Plan()
        lp.sb0 = getattr(plan, 'sb0', None)
        lp.adjb = getattr(plan, 'adjb', None)
        lp.asym = getattr(plan, 'asym', None)
        lp.adj_flags = getattr(plan, 'adj_flags', None)
        lp.S_full = getattr(plan, 'S0_full', None)
        lp.enc_plan = plan
        lp.ce_scale = self._ce_scale
        lp.ent_w = ent_w
        lp.link_kind = self.link_loss_kind if self.linkpred else None
        lp.packed = bool(getattr(plan, 'packed', False)) and torch.is_tensor(adj) and adj.dtype == torch.float32
        N0 = S0.shape[1]
        if batch_num_nodes is None and adj is None:
            lp.nb_dev, nb_host = plan.nb_dev, plan.nb_host                   # 2-argument call: the forward's n_b
        else:
            lp.nb_dev, nb_host = E.prep_nb(batch_num_nodes, N0, S0.device)
        lp.inv_dev = None
        dev_only = nb_host is None and lp.nb_dev is not None               # node counts on the device only
        if dev_only and ent_w != 0.0:
            raise NotImplementedError('gp_b200: entropy_weight with device-resident batch_num_nodes')
        lp.num_real_rows = S0.shape[0] * N0 if nb_host is None else max(int(np.sum(nb_host.astype(np.int64))), 1)
        if self.linkpred:
            if not isinstance(adj, T.PreparedAdjacency):
                adj = E._chk_adj(adj, lp.sb0 is not None)
            if dev_only:
                st = torch.empty(2, device=S0.device, dtype=torch.float32)
                call('gp_nb_stats', lp.nb_dev.data_ptr(), S0.shape[0], st.data_ptr(), E._stream())
                lp.inv_dev, lp.num_entries = st[0:1], 1
            elif nb_host is None:
                lp.num_entries = adj.shape[1] * adj.shape[1] * adj.shape[0]
                print('Warning: calculating link pred loss without masking')       # encoders.py:1324
            else:
                n64 = nb_host.astype(np.int64)                                   # R11
                lp.num_entries = int(np.sum(n64 * n64))
            if self._entries_override is not None:
                lp.num_entries = self._entries_override
                if dev_only:        # the override (dp.py, mode='global_norm') replaces the device-side 1 / sum n_b^2 too
                    call('gp_fill_f32', lp.inv_dev.data_ptr(), C.c_longlong(1), C.c_float(1.0 / float(lp.num_entries)),
                         E._stream())
        if self.linkpred and adj_hop > 1:
            # encoders.py:1312-1317, never passed by the reference's callers: fp32 FFMA schedule, fp32 dense adjacency
            if self.link_loss_kind != 'bce' or dev_only or isinstance(adj, T.PreparedAdjacency) or \
                    adj.dtype != torch.float32:
                raise NotImplementedError('gp_b200: adj_hop > 1 needs the BCE link loss, a float32 adjacency tensor and '
                                          'host-side batch_num_nodes')
            lp.link_kind = None
            outs = _LossFn.apply(lp, pred, label, S0, None)
            total, self.link_loss = _LinkHopFn.apply(S0, outs[0], adj, lp.nb_dev, 1.0 / float(lp.num_entries), adj_hop)
            if ent_w != 0.0:
                self.entropy_loss = outs[1]
            return total
        outs = _LossFn.apply(lp, pred, label, S0, adj if self.linkpred else None)   # adj: tensor or PreparedAdjacency
        k = 1
        if self.linkpred:
            self.link_loss = outs[k]
            k += 1
        if ent_w != 0.0:
            self.entropy_loss = outs[k]
        return outs[0]
