"""Per-call timing of ONE training step through the C ABI, for bench.py's `roofline.kernels` table.

While a CallProfiler is installed (`_lib.set_hook`), every `_lib.call` is bracketed by CUDA events on the launching
stream; after the step the calls are grouped by (entry point, shape) and each group gets its ALGORITHMIC work --
flops and compulsory bytes derived from the call's own arguments (dense extents; the dtype the bytes are counted at
is the dtype that call reads / writes and is stated per row) -- the roof that bounds it at the measured peaks, and
the fraction of that roof it achieved.  The events serialise nothing that was not already serial (one stream), and
the profiled step is an extra, untimed one: no number measured here is a bench value.
"""
import ctypes as C

import torch

from . import _lib


def _ival(a):
    if isinstance(a, int):
        return a
    v = getattr(a, 'value', a)
    return 0 if v is None else v


class CallProfiler:
    def __init__(self):
        self.rows = []

    # ---- hook protocol -------------------------------------------------------------------------------
    def begin(self, name, args):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (name, self.describe(name, args), e0)

    def end(self, tok):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.rows.append(tok + (e1,))

    def __enter__(self):
        _lib.set_hook(self)
        return self

    def __exit__(self, *exc):
        _lib.set_hook(None)

    # ---- work models ---------------------------------------------------------------------------------
    @staticmethod
    def describe(name, args):
        """-> (shape key, algorithmic flops, algorithmic bytes, dtype note) for the entry points that matter."""
        try:
            if name in ('gp_bgemm_bf16x', 'gp_bgemm_bf16_norm'):
                g = args[0]._obj
                M, N, batch = g.M, g.N, g.batch
                ks, fl, by = [], 0.0, 0.0
                for q in range(g.npairs):
                    pr = g.pair[q]
                    ks.append(pr.K)
                    fl += 2.0 * M * N * pr.K * batch
                    by += 2.0 * (M * pr.K * (batch if pr.sAb or batch == 1 else 1) + pr.K * N * (batch if pr.sBb or batch == 1 else 1))
                if g.C:
                    by += 4.0 * M * N * batch * (2 if g.beta != 0.0 else 1)
                if g.Cb:
                    by += 2.0 * M * N * batch
                tag = 'tail' if name.endswith('norm') else 'gemm'
                key = '%s M=%d N=%d K=%s batch=%d%s' % (tag, M, N, '+'.join(map(str, ks)), batch,
                                                       ' split_k=%d' % g.split_k if g.split_k > 1 else '')
                return key, fl, by, 'bf16 operands, fp32/bf16 outputs'
            if name == 'gp_linkloss_tc':
                B, N, K = _ival(args[5]), _ival(args[6]), _ival(args[7])
                has_g = bool(_ival(args[9]))
                f, tag = 1.0, ''
                if _ival(args[11]) == 2:        # symmetric {0,1} adjacency: only the upper diagonal band is computed / written
                    tm, tn = (N + 127) // 128, (N + 255) // 256
                    f = sum(max(0, tn - (mt * 128) // 256) for mt in range(tm)) / float(tm * tn)
                    tag = ' upper band (%.0f %% of the tiles, if the batch is symmetric {0,1})' % (100 * f)
                return ('linkloss fwd N=%d K=%d batch=%d%s' % (N, K, B, tag), 2.0 * N * N * K * B * f,
                        2.0 * B * (N * K + f * N * N + (f * N * N if has_g else 0)), 'bf16 S, bf16 adjacency, bf16 G')
            if name == 'gp_pool_chain_bf16':
                B, N, K = _ival(args[6]), _ival(args[7]), _ival(args[8])
                has_t = bool(_ival(args[9]))
                by = B * (2.0 * N * K + 2.0 * N * N + (2.0 * K * N if has_t else 0.0) +
                          (4.0 if _ival(args[11]) else 0.0) * K * K + (2.0 if _ival(args[13]) else 0.0) * K * K)
                return ('chained S^T A S N=%d K=%d batch=%d%s' % (N, K, B, ' (+T store)' if has_t else ''),
                        2.0 * B * (float(K) * N * N + float(K) * K * N), by,
                        'bf16 S and adjacency in, fp32 + bf16 A\' out, T on chip' + (' and stored once as bf16' if has_t else ''))
            if name in ('gp_adj_prepare', 'gp_adj_prepare_x'):
                off = 1 if name.endswith('_x') else 0
                kind = _ival(args[1])
                B, N = _ival(args[3 + off]), _ival(args[4 + off])
                rd = {0: 4.0, 1: 1.0, 2: 0.125}[kind]
                return ('adj_prepare kind=%d N=%d batch=%d' % (kind, N, B), 0.0, B * N * N * (rd + 2.0),
                        '%s adjacency in, bf16 operand out' % {0: 'fp32', 1: 'uint8', 2: 'bit-packed'}[kind])
            if name == 'gp_gcn_layer_bwd_x':
                q = args[0]._obj
                rows = float(q.B) * q.N
                src_b = (0.0 if not q.dz else (2.0 if q.dz_bf16 else 4.0)) + (0.0 if not q.dxn else (2.0 if q.dxn_bf16 else 4.0))
                by = rows * q.d * (src_b + 4.0 + (2.0 if q.dv_bf16 else 0.0) + (4.0 if q.dv else 0.0))
                return ('layer_bwd d=%d bn=%d rows=%d%s' % (q.d, q.bn, int(rows), ' (bf16 gradient sources)' if (q.dz_bf16 or q.dxn_bf16) else ''),
                        0.0, by, 'fp32 Y + fp32 / bf16 gradient sources in, bf16/fp32 dV out')
            if name == 'gp_bn_apply':
                B, N, d = _ival(args[4]), _ival(args[5]), _ival(args[6])
                outs = (4.0 if _ival(args[9]) else 0.0) + (2.0 if _ival(args[11]) else 0.0) + (2.0 if _ival(args[13]) else 0.0)
                return ('bn_apply d=%d rows=%d' % (d, B * N), 0.0, float(B) * N * d * (4.0 + outs), 'fp32 in, fp32 (embedding stack only) + bf16 out')
            if name in ('gp_softmax_mask_fwd_x', 'gp_softmax_mask_bwd_x'):
                o = 2 if name.endswith('fwd_x') else 3
                B, N, K = _ival(args[o]), _ival(args[o + 1]), _ival(args[o + 2])
                per = 10.0 if name.endswith('fwd_x') else 10.0          # fwd: read+write fp32 + bf16 ; bwd: 2 fp32 in, bf16 out
                return ('%s K=%d rows=%d' % (name[3:], K, B * N), 0.0, float(B) * N * K * per, 'fp32 rows, bf16 copy')
        except Exception:                       # a model must never break a profiled step
            pass
        return name[3:], None, None, ''

    # ---- aggregation ---------------------------------------------------------------------------------
    def table(self, hbm_gbs, tf_peak, min_share=0.03):
        """Group by (entry point, shape).  Returns (rows sorted by time, total ms of the profiled calls)."""
        torch.cuda.synchronize()
        agg = {}
        for name, (key, fl, by, note), e0, e1 in self.rows:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault((name, key), dict(entry=name, shape=key, launches=0, ms=0.0, flops=0.0, bytes=0.0,
                                                 counted_at=note, modelled=fl is not None))
            a['launches'] += 1
            a['ms'] += ms
            if fl is not None:
                a['flops'] += fl
                a['bytes'] += by
        total = sum(a['ms'] for a in agg.values())
        out = []
        for a in sorted(agg.values(), key=lambda r: -r['ms']):
            a['share'] = a['ms'] / total if total else 0.0
            if a['share'] < min_share:
                continue
            if a['modelled'] and a['ms'] > 0:
                t = a['ms'] * 1e-3
                t_tc, t_hbm = a['flops'] / (tf_peak * 1e12), a['bytes'] / (hbm_gbs * 1e9)
                if t_tc > t_hbm:
                    a.update(bound='tensor', achieved=a['flops'] / t / 1e12, peak=tf_peak, unit='TFLOP/s')
                else:
                    a.update(bound='hbm', achieved=a['bytes'] / t / 1e9, peak=hbm_gbs, unit='GB/s')
                a['frac'] = a['achieved'] / a['peak']
            a['ms_per_launch'] = a['ms'] / a['launches']
            out.append(a)
        return out, total
