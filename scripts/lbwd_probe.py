"""Times gp_gcn_layer_bwd_x (layer tail backward) alone at a given shape.  usage: python scripts/lbwd_probe.py [B N d [bn]]
(bn = 0: the last layer of a stack -- no ReLU / BatchNorm, rows independent -- e.g. 256 2048 512 0)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_pooling_b200 import engine as E, engine_tc as T
from graph_pooling_b200._lib import GpLayerBwd, call, load
B, N, d = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (256, 2048, 128)
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 1
h16 = len(sys.argv) > 5 and sys.argv[5] == 'bf16'      # gradient sources as bf16 buffers
dev = torch.device('cuda')
ws = E.Workspace(dev)
Fw = 3 * d if bn else d + 256
dz, dxn = torch.randn(B, N, Fw, device=dev), torch.randn(B, N, d, device=dev)
if h16:
    dz, dxn = dz.bfloat16(), dxn.bfloat16()
y, rn = torch.randn(B, N, d, device=dev), torch.rand(B, N, device=dev) + 0.5
mean, invstd = torch.randn(N, device=dev) * 0.1, torch.rand(N, device=dev) + 0.5
dvb = T.bfbuf(ws, 1, B * N, d)
db = ws.f(d)
q = GpLayerBwd()
q.dz, q.lddz, q.dxn, q.dout, q.argidx, q.ldo = dz.data_ptr(), Fw, (dxn.data_ptr() if bn else None), None, None, 0
q.h, q.ldh, q.y, q.ldy = None, Fw, y.data_ptr(), d
q.rnorm, q.mean, q.invstd = rn.data_ptr(), mean.data_ptr(), invstd.data_ptr()
q.B, q.N, q.d, q.relu, q.bn, q.normalize = B, N, d, bn, bn, 1
q.dv, q.dv_bf16, q.lddvb, q.db = None, dvb.ptr, dvb.ld, db.data_ptr()
q.dz_bf16, q.dxn_bf16 = int(h16), int(h16 and bn)
q.ws = None
w = ws.f(int(load().gp_gcn_layer_bwd_ws_x(C.byref(q))))
q.ws = w.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    call('gp_gcn_layer_bwd_x', C.byref(q), st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    call('gp_gcn_layer_bwd_x', C.byref(q), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
gb = B * N * d * ((2 if h16 else 4) * (2 if bn else 1) + 4 + 2) / 1e9
print('layer_bwd_x B=%d N=%d d=%d bn=%d VPT=%s: %.3f ms  %.0f GB/s (algorithmic %.2f GB)' % (B, N, d, bn, os.environ.get('GP_LBWD_VPT', '8'), ms, gb / ms * 1e3, gb))
