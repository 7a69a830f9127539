"""ctypes binding of libgp_b200.so (the C ABI declared in include/gp_b200.h).

The product path has NO fallback: if the library is missing or a call fails, we raise.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libgp_b200.so')

c_f = C.c_void_p      # device pointers travel as integers
c_ll = C.c_longlong
c_i = C.c_int


class GpGemm(C.Structure):
    _fields_ = [('A', c_f), ('B', c_f), ('C', c_f),
                ('M', c_i), ('N', c_i), ('K', c_i), ('batch', c_i),
                ('sAb', c_ll), ('sAm', c_ll), ('sAk', c_ll),
                ('sBb', c_ll), ('sBk', c_ll), ('sBn', c_ll),
                ('sCb', c_ll), ('sCm', c_ll), ('sCn', c_ll),
                ('lim', c_f), ('lim_m', c_i), ('lim_n', c_i), ('lim_k', c_i),
                ('alpha', C.c_float), ('beta', C.c_float),
                ('alpha_dev', c_f), ('bias', c_f), ('relu', c_i), ('split_k', c_i)]


class GpGemmBf16(C.Structure):
    _fields_ = [('A', c_f), ('B', c_f), ('C', c_f), ('Cb', c_f),
                ('M', c_i), ('N', c_i), ('K', c_i), ('batch', c_i),
                ('ldA', c_ll), ('sAb', c_ll), ('a_major', c_i),
                ('ldB', c_ll), ('sBb', c_ll), ('b_major', c_i),
                ('ldC', c_ll), ('sCb', c_ll), ('ldCb', c_ll), ('sCbb', c_ll),
                ('lim', c_f), ('lim_m', c_i), ('lim_n', c_i), ('lim_k', c_i),
                ('alpha', C.c_float), ('beta', C.c_float), ('alpha_dev', c_f),
                ('bias', c_f), ('relu', c_i), ('split_k', c_i)]


class GpOperandPair(C.Structure):
    _fields_ = [('A', c_f), ('B', c_f), ('K', c_i),
                ('ldA', c_ll), ('sAb', c_ll), ('a_major', c_i),
                ('ldB', c_ll), ('sBb', c_ll), ('b_major', c_i), ('lim_k', c_i)]


class GpGemmBf16x(C.Structure):
    _fields_ = [('pair', GpOperandPair * 4), ('npairs', c_i),
                ('C', c_f), ('Cb', c_f),
                ('M', c_i), ('N', c_i), ('batch', c_i),
                ('ldC', c_ll), ('sCb', c_ll), ('ldCb', c_ll), ('sCbb', c_ll),
                ('lim', c_f), ('lim_m', c_i), ('lim_n', c_i),
                ('alpha', C.c_float), ('beta', C.c_float), ('alpha_dev', c_f),
                ('bias', c_f), ('relu', c_i), ('split_k', c_i),
                ('cond', c_f), ('cond_npairs', c_i), ('cond_alpha', C.c_float), ('order', c_f), ('tri', c_i)]


class GpAxpyEntry(C.Structure):
    _fields_ = [('src', c_f), ('dst', c_f), ('n', c_ll)]


class GpLayerBwd(C.Structure):
    _fields_ = [('dz', c_f), ('lddz', c_ll), ('dxn', c_f), ('dout', c_f), ('argidx', c_f), ('ldo', c_ll),
                ('h', c_f), ('ldh', c_ll), ('y', c_f), ('ldy', c_ll),
                ('rnorm', c_f), ('mean', c_f), ('invstd', c_f),
                ('B', c_i), ('N', c_i), ('d', c_i), ('relu', c_i), ('bn', c_i), ('normalize', c_i),
                ('dv', c_f), ('dv_bf16', c_f), ('lddvb', c_ll), ('db', c_f), ('ws', c_f), ('lddxn', c_ll),
                ('nb_zero', c_f), ('dz_bf16', c_i), ('dxn_bf16', c_i)]


# ---- packed small-graph schedule (include/gp_b200.h "PACKED schedule") ---------------------------------------------
PK_MAX_LAYERS = 6


PK_ELL = 8


class GpPkTiling(C.Structure):
    _fields_ = [('rowptr', c_f), ('subs', c_f), ('nsub', c_f), ('rowmeta', c_f), ('B', c_i), ('nfix', c_i),
                ('gpt', c_i), ('max_rows', c_i)]


class GpPkAdj(C.Structure):
    _fields_ = [('info', c_f), ('ell', c_f), ('entries', c_f), ('dense', c_f), ('transposed', c_i)]


class GpPkSrc(C.Structure):
    _fields_ = [('y', c_f), ('ld', c_ll), ('d', c_i), ('sums', c_f), ('bias', c_f)]


class GpPkGrad(C.Structure):
    _fields_ = [('dense', c_f), ('ld', c_ll), ('coff', c_i), ('dout', c_f), ('arg', c_f), ('ldo', c_ll), ('ooff', c_i)]


class GpPkStackFwd(C.Structure):
    _fields_ = [('inp', GpPkSrc), ('W', c_f), ('b', c_f), ('dout', c_i), ('y', c_f), ('rnorm', c_f), ('sums_out', c_f)]


class GpPkLayerFwdArgs(C.Structure):
    _fields_ = [('tl', GpPkTiling), ('adj', GpPkAdj), ('cnt_pad', c_f), ('N', c_i), ('ns', c_i), ('s', GpPkStackFwd * 2)]


class GpPkStackBwd(C.Structure):
    _fields_ = [('inp', GpPkSrc), ('out', GpPkSrc), ('rnorm', c_f), ('msums', c_f), ('gl', GpPkGrad),
                ('W', c_f), ('b', c_f), ('dout', c_i), ('dW', c_f), ('db', c_f), ('need_dx', c_i),
                ('gz_prev', GpPkGrad), ('gl_prev', c_f), ('msums_prev', c_f), ('dadj', c_f), ('dadj_acc', c_i)]


class GpPkLayerBwdArgs(C.Structure):
    _fields_ = [('tl', GpPkTiling), ('adj', GpPkAdj), ('adj_in', GpPkAdj), ('cnt_pad', c_f), ('N', c_i), ('ns', c_i),
                ('s', GpPkStackBwd * 2)]


class GpPkConcat(C.Structure):
    _fields_ = [('L', c_i), ('F', c_i), ('slot', GpPkSrc * PK_MAX_LAYERS)]


class GpPkPoolArgs(C.Structure):
    _fields_ = [('tl', GpPkTiling), ('adj', GpPkAdj), ('adj_in', GpPkAdj), ('cnt_pad', c_f), ('nb', c_f), ('N', c_i),
                ('K', c_i), ('z', GpPkConcat), ('za', GpPkConcat), ('Wp', c_f), ('bp', c_f), ('S', c_f), ('xp', c_f),
                ('ap', c_f), ('out', c_f), ('arg', c_f), ('ldo', c_ll), ('dxp', c_f), ('dap', c_f), ('dS_ext', c_f),
                ('dout', c_f), ('gz', c_f), ('gza', c_f), ('dWp', c_f), ('dbp', c_f)]


# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    'gp_bgemm_bf16x': [C.POINTER(GpGemmBf16x), c_f],
    'gp_bgemm_bf16': [C.POINTER(GpGemmBf16), c_f],
    'gp_bgemm_bf16_norm': [C.POINTER(GpGemmBf16x), c_f, c_f, c_i, c_f],
    'gp_bn_finalize': [c_f, c_i, c_i, c_i, c_f, c_f, c_f],
    'gp_bn_apply': [c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f],
    'gp_bias_normalize_x': [c_f, c_f, c_f, c_ll, c_i, c_ll, c_i, c_f, c_ll, c_f],
    'gp_softmax_mask_fwd_x': [c_f, c_f, c_i, c_i, c_i, c_f, c_ll, c_f],
    'gp_softmax_mask_bwd_x': [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_ll, c_f, c_f, c_f],
    'gp_adj_prepare': [c_f, c_i, c_f, c_i, c_i, c_f, c_ll, c_f, c_f],
    'gp_adj_prepare_x': [c_f, c_i, c_ll, c_f, c_i, c_i, c_f, c_ll, c_f, c_i, c_f],
    'gp_adj_from_edges': [c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_f, c_ll, c_f, c_i, c_f],
    'gp_host_pack_adj_bits': [c_f, c_ll, c_i, c_f, c_ll, c_i, c_f],
    'gp_sym_select_bf16': [c_f, c_ll, c_i, c_i, c_f, c_f, c_ll, c_f],
    'gp_cvt_f32_bf16': [c_f, c_ll, c_f, c_ll, c_ll, c_i, c_i, c_f],
    'gp_version': [],
    'gp_last_error': [],
    'gp_launch_count': [],
    'gp_launch_count_reset': [],
    'gp_bgemm_f32': [C.POINTER(GpGemm), c_f],
    'gp_graphconv_fwd': [c_f, c_ll, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_ll, c_f, c_i, c_f],
    'gp_graphconv_bwd': [c_f, c_f, c_f, c_ll, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f,
                         c_i, c_f],
    'gp_graphconv_bwd_ws': [c_i, c_i, c_i, c_i, c_i],
    'gp_relu_bn_fwd': [c_f, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f],
    'gp_relu_bn_fwd_ws': [c_i, c_i, c_i],
    'gp_relu_bn_fwd_x': [c_f, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f],
    'gp_gcn_layer_bwd': [c_f, c_ll, c_f, c_f, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_i,
                         c_i, c_f, c_f],
    'gp_gcn_layer_bwd_x': [C.POINTER(GpLayerBwd), c_f],
    'gp_gcn_layer_bwd_ws': [c_i, c_i, c_i, c_i],
    'gp_gcn_layer_bwd_ws_x': [C.POINTER(GpLayerBwd)],
    'gp_readout_max_fwd': [c_f, c_ll, c_f, c_i, c_i, c_i, c_f, c_f, c_ll, c_f],
    'gp_readout_max_fwd_x': [c_f, c_ll, c_i, c_f, c_i, c_i, c_i, c_f, c_f, c_ll, c_f],
    'gp_softmax_mask_fwd': [c_f, c_f, c_i, c_i, c_i, c_f],
    'gp_softmax_mask_bwd': [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f],
    'gp_pool_fwd': [c_f, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_i, c_f],
    'gp_pool_bwd': [c_f, c_f, c_f, c_f, c_ll, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_ll, c_i, c_f, c_i, c_f,
                    c_f, c_i, c_f],
    'gp_linkloss_fwd': [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f],
    'gp_loss_finalize': [c_f, c_i, C.c_double, c_f, c_f, c_f, c_f],
    'gp_linkloss_tc': [c_f, c_ll, c_f, c_ll, c_f, c_i, c_i, c_i, c_f, c_f, c_ll, c_i, c_f, c_f],
    'gp_frob_link_fwd': [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f],
    'gp_frob_finalize': [c_f, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f],
    'gp_scale_rows_batch': [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_ll, c_f, c_ll, c_i, c_f],
    'gp_entropy_partials': [c_i, c_i],
    'gp_entropy_fwd': [c_f, c_f, c_i, c_i, c_i, c_f, c_f],
    'gp_entropy_bwd': [c_f, c_f, c_i, c_i, c_i, c_f, C.c_float, c_f, c_i, c_f],
    'gp_sumsq_f32': [c_f, c_ll, c_f, c_f, c_f],
    'gp_adam_step_f32': [c_f, c_f, c_f, c_f, c_ll, C.c_float, C.c_float, C.c_float, C.c_float, c_f, c_f, C.c_float,
                         C.c_float, c_f],
    'gp_clip_scale_f32': [c_f, c_ll, c_f, C.c_float, C.c_float, c_f],
    'gp_multi_axpy_f32': [c_f, c_i, C.c_float, c_f],
    'gp_nb_stats': [c_f, c_i, c_f, c_f],
    'gp_mul_add_dev': [c_f, c_f, c_f, c_f, c_f, c_f],
    'gp_add_scaled': [c_f, c_f, C.c_float, c_f, c_f],
    'gp_linkloss_tc_partials': [c_i, c_i],
    'gp_gcn_layer_bwd_vectorised': [c_i, c_i, c_i],
    'gp_gcn_layer_bwd_bf16_sources_fast': [c_i, c_i, c_i],
    'gp_pool_chain_bf16': [c_f, c_ll, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f],
    'gp_linkloss_from_q_partials': [c_i, c_i],
    'gp_linkloss_from_q': [c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f],
    'gp_ce_fwd': [c_f, c_f, c_i, c_i, c_f, c_f, c_f],
    'gp_ce_bwd': [c_f, c_f, c_f, c_i, c_i, c_f, c_f],
    'gp_colsum_f32': [c_f, c_ll, c_i, c_ll, c_f, c_i, c_f, c_f],
    'gp_relu_mask_bwd': [c_f, c_f, c_ll, c_f, c_f],
    'gp_bias_normalize_f32': [c_f, c_f, c_f, c_ll, c_i, c_ll, c_i, c_f],
    'gp_fill_f32': [c_f, c_ll, C.c_float, c_f],
    'gp_axpy_f32': [c_f, c_f, c_ll, C.c_float, c_f],
    'gp_fill_i32': [c_f, c_ll, c_i, c_f],
    'gp_dropout_f32': [c_f, c_ll, c_ll, c_i, C.c_float, C.c_ulonglong, c_f, c_ll, c_f, c_ll, c_f],
    'gp_set2set_fwd': [c_f, c_ll, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f],
    'gp_set2set_bwd': [c_f, c_ll, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_ll, c_f, c_f, c_f, c_f],
    'gp_pk_prepare': [c_f, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f],
    'gp_pk_build_lists': [c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_ll, c_f, c_f, c_i, c_f, c_ll, c_f,
                          c_i, c_f, c_ll, c_f],
    'gp_pk_layer_fwd': [C.POINTER(GpPkLayerFwdArgs), c_f],
    'gp_pk_layer_bwd': [C.POINTER(GpPkLayerBwdArgs), c_f],
    'gp_pk_pool_fwd': [C.POINTER(GpPkPoolArgs), c_f],
    'gp_pk_pool_bwd': [C.POINTER(GpPkPoolArgs), c_f],
    'gp_pk_readout': [C.POINTER(GpPkTiling), C.POINTER(GpPkConcat), c_f, c_f, c_i, c_f, c_f, c_ll, c_i, c_f],
    'gp_pk_link_fwd': [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f],
    'gp_pk_link_bwd': [c_f, c_f, c_f, c_i, c_i, c_i, C.c_float, c_f, c_f, c_f, c_f],
    'gp_pk_link_finalize': [c_f, C.c_double, c_f, c_f, c_f, c_f, c_f],
    'gp_pad_copy_f32': [c_f, c_ll, c_ll, c_i, c_f, c_ll, c_ll, c_i, C.c_float, c_f],
}
_RESTYPES = {'gp_last_error': C.c_char_p, 'gp_launch_count': c_ll, 'gp_gcn_layer_bwd_ws': c_ll, 'gp_gcn_layer_bwd_ws_x': c_ll, 'gp_relu_bn_fwd_ws': c_ll, 'gp_graphconv_bwd_ws': c_ll, 'gp_launch_count_reset': None}

EXPORTS = tuple(_PROTOS)
_lib = None


class GpError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpError('libgp_b200.so not found at %s -- run `python -c "import __graft_entry__ as g; g.build()"` '
                      '(there is no CPU or PyTorch fallback for the hot path)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, c_i)
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        raise GpError('%s failed (%d): %s' % (what, status, load().gp_last_error().decode()))


_hook = None     # profile.CallProfiler while a profiled step runs (bench.py); None on the product path


def set_hook(h):
    global _hook
    _hook = h


class _NvtxHook:
    """GP_NVTX=1: every C-ABI call (= one fused op of the hot path) becomes an NVTX range named after its entry point,
    so Nsight timelines / `ncu --nvtx --nvtx-include` can address the ops by name.  Off by default (no overhead)."""

    def begin(self, name, args):
        import torch
        torch.cuda.nvtx.range_push(name)
        return name

    def end(self, tok):
        import torch
        torch.cuda.nvtx.range_pop()


if os.environ.get('GP_NVTX'):
    _hook = _NvtxHook()


def call(name, *args):
    if _hook is None:
        check(getattr(load(), name)(*args), name)
        return
    tok = _hook.begin(name, args)
    check(getattr(load(), name)(*args), name)
    _hook.end(tok)
