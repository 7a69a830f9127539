"""T = S^T.A on a ragged batch (cfg5 shapes): effect of the cluster width K being a multiple of 8 / 128 or not."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_pooling_b200 import engine as E, engine_tc as T
dev = torch.device('cuda'); ws = E.Workspace(dev)
B, N = 32, 5000
rs = np.random.RandomState(0)
nb = rs.randint(50, N + 1, size=B).astype(np.int32)
nbd = torch.tensor(nb).cuda()
adjb = T.bfbuf(ws, B, N, N); adjb.t.zero_()
def timeit(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for K in (1250, 1256, 1280):
    sb = T.bfbuf(ws, B, N, K); sb.t.normal_()
    tb = T.bfbuf(ws, B, K, N)
    fl = float(np.sum(2.0 * K * nb.astype(np.float64) ** 2))
    t = timeit(lambda: T.tcgemm(sb, T.MN, adjb, T.MN, K, N, N, B, Cb=tb, lim=nbd.data_ptr(), lim_k=1, lim_n=1))
    t2 = timeit(lambda: T.tcgemm(sb, T.MN, adjb, T.MN, K, N, N, B, Cb=tb))
    print('K=%d: ragged %.3f ms (%.0f TFLOP/s useful)   dense(no skip) %.3f ms (%.0f TFLOP/s)' % (K, t, fl / t / 1e9, t2, 2.0 * K * N * N * B / t2 / 1e9))
