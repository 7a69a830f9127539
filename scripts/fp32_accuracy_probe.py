"""Where does the fp32 schedule lose accuracy at N = 5000 / K = 1250?  Each op of the backward at that size, on
operands with the magnitudes of the real step, against the same formula in fp64 (torch on the GPU).  Diagnostic."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_pooling_b200 import engine as E  # noqa: E402
from graph_pooling_b200._lib import call  # noqa: E402


def rel(a, b):
    return float((a.double() - b).norm() / b.norm().clamp_min(1e-300))


def main():
    N, K, H, F = int(os.environ.get('PN', 5000)), int(os.environ.get('PK', 1250)), 128, 384
    B = 2
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(0)
    nb = torch.tensor([N, max(N // 100, 8)], dtype=torch.int32, device=dev)
    real = (torch.arange(N, device=dev)[None, :] < nb[:, None]).float()
    u = torch.triu(torch.rand(B, N, N, device=dev, generator=g) < 8.0 / N, 1).float()
    A = (u + u.transpose(1, 2)) * real[:, :, None] * real[:, None, :]
    S = torch.softmax(0.3 * torch.randn(B, N, K, device=dev, generator=g), -1) * real[:, :, None]
    Z = torch.randn(B, N, F, device=dev, generator=g) * 0.1 * real[:, :, None]
    ws = E.Workspace(dev)
    st = E._stream()
    A64, S64, Z64 = A.double(), S.double(), Z.double()
    nbp = nb.data_ptr()
    print('N=%d K=%d' % (N, K))
    # 1. U = A.X
    X = torch.randn(B, N, H, device=dev, generator=g)
    U = ws.f(B, N, H)
    E.bgemm(A.data_ptr(), X.data_ptr(), U.data_ptr(), N, H, N, B, (N * N, N, 1), (N * H, H, 1), (N * H, H, 1), lim=nbp,
            lim_m=1, lim_k=1)
    print('U = A.X                ', rel(U, A64 @ X.double()))
    # 2. pool forward
    xp, t, ap = ws.f(B, K, F), ws.f(B, K, N), ws.f(B, K, K)
    call('gp_pool_fwd', S.data_ptr(), Z.data_ptr(), F, A.data_ptr(), nbp, B, N, K, F, xp.data_ptr(), t.data_ptr(),
         ap.data_ptr(), 0, st)
    T64 = S64.transpose(1, 2) @ A64
    print("X' = S^T Z             ", rel(xp, S64.transpose(1, 2) @ Z64))
    print('T  = S^T A             ', rel(t, T64))
    print("A' = T S               ", rel(ap, T64 @ S64))
    # 3. pool backward
    dxp = torch.randn(B, K, F, device=dev, generator=g) * 1e-3
    dap = torch.randn(B, K, K, device=dev, generator=g) * 1e-3
    dz, ds, wsp = ws.f(B, N, F), ws.f(B, N, K), ws.f(B, N, K)
    call('gp_pool_bwd', dxp.data_ptr(), dap.data_ptr(), S.data_ptr(), Z.data_ptr(), F, A.data_ptr(), t.data_ptr(), nbp,
         B, N, K, F, dz.data_ptr(), F, 0, ds.data_ptr(), 0, None, wsp.data_ptr(), 0, st)
    dap64, dxp64 = dap.double(), dxp.double()
    ds_pool64 = Z64 @ dxp64.transpose(1, 2) + T64.transpose(1, 2) @ dap64 + A64 @ (S64 @ dap64.transpose(1, 2))
    ds_pool64 = ds_pool64 * real[:, :, None].double()
    print('dZ = S dX\'             ', rel(dz, S64 @ dxp64))
    print('dS (pool)              ', rel(ds, ds_pool64))
    # 4. link loss forward / backward
    nt = (N + 63) // 64
    partial, gsym = ws.f(B * nt * nt + 256), ws.f(B, N, N)
    call('gp_linkloss_fwd', S.data_ptr(), A.data_ptr(), nbp, B, N, K, partial.data_ptr(), gsym.data_ptr(), st)
    P64 = S64 @ S64.transpose(1, 2)
    m2 = (real[:, :, None] * real[:, None, :]).double()
    G64 = (-A64 / (P64 + 1e-7) + (1 - A64) / (1 - P64 + 1e-7)) * m2
    print('gsym = G + G^T         ', rel(gsym, G64 + G64.transpose(1, 2)))
    inv = 1.0 / float((nb.double() ** 2).sum())
    one = torch.ones(1, device=dev)
    dsl = ws.f(B, N, K)
    E.bgemm(gsym.data_ptr(), S.data_ptr(), dsl.data_ptr(), N, K, N, B, (N * N, N, 1), (N * K, K, 1), (N * K, K, 1),
            lim=nbp, lim_m=1, lim_k=1, alpha=inv, alpha_dev=one.data_ptr())
    ds_link64 = (G64 + G64.transpose(1, 2)) @ S64 * inv
    print('dS (link) = gsym.S     ', rel(dsl, ds_link64), ' |dS_link|/|dS_pool| = %.3g' % float(ds_link64.norm() / ds_pool64.norm()))
    # 5. softmax backward on the exact fp64 dS (rounded to fp32): conditioning of the projection
    ds_tot64 = ds_pool64 + ds_link64
    ds_tot = ds_tot64.float().contiguous()
    dt = ws.f(B, N, K)
    call('gp_softmax_mask_bwd', S.data_ptr(), ds_tot.data_ptr(), nbp, B, N, K, dt.data_ptr(), st)
    dt64 = S64 * (ds_tot64 - (ds_tot64 * S64).sum(-1, keepdim=True))
    dt_in32 = S64 * (ds_tot.double() - (ds_tot.double() * S64).sum(-1, keepdim=True))
    print('dT softmax bwd (kernel vs fp64 on the same fp32 input)', rel(dt, dt_in32))
    print('dT: effect of rounding dS to fp32 alone               ', rel(dt_in32.float(), dt64),
          ' amplification |dS|/|dS - <dS,S>| = %.3g' % float(ds_tot64.norm() / (ds_tot64 - (ds_tot64 * S64).sum(-1, keepdim=True)).norm()))
    # the same with the dS the kernels produced
    ds_k = (ds + dsl).contiguous()
    call('gp_softmax_mask_bwd', S.data_ptr(), ds_k.data_ptr(), nbp, B, N, K, dt.data_ptr(), st)
    print('dT from the kernels\' dS                               ', rel(dt, dt64))
    # 6. dWp = dT^T za (split-K over B*N rows)
    Fa = 2 * H + K
    za = torch.randn(B * N, Fa, device=dev, generator=g) * 0.1
    dtf = dt64.float().reshape(B * N, K).contiguous()
    dwp = ws.f(K, Fa)
    split = max(1, min(512, (B * N) // 1024))
    E.bgemm(dtf.data_ptr(), za.data_ptr(), dwp.data_ptr(), K, Fa, B * N, 1, (0, 1, K), (0, Fa, 1), (0, Fa, 1), split_k=split)
    print('dWp = dT^T za (split %d) ' % split, rel(dwp, dtf.double().t() @ za.double()))
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
