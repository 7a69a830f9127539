"""CPU: the algorithmic FLOP / byte model behind bench.py's roofline numbers (graph_pooling_b200/roofline.py) against the
figures SURVEY.md 8(d) derives for the named configurations, and against a brute-force count of the products the
reference's forward / backward performs (encoders.py:1231-1300 with the association (A.X).W, no dA at level 0)."""
import numpy as np

from graph_pooling_b200 import roofline, synth


def test_survey_figures():
    cfg4 = dict(synth.WORKLOADS['cfg4_diffpool_256x2048'])
    f, b = roofline.graph_flops(2048, cfg4)
    assert abs((f + b) / 1e9 - 45.75) < 0.01                       # 45.75 GFLOP per graph
    assert abs(256 * (f + b) / 1e12 - 11.71) < 0.01                # 11.7 TFLOP per 256-graph step
    assert abs(roofline.step_bytes([2048] * 256, cfg4) / 1e9 - 16.1) < 0.05
    cfg1 = dict(synth.WORKLOADS['cfg1_enzymes_like'])
    f, b = roofline.graph_flops(32.2, cfg1)
    assert 1.8e6 < f + b < 2.3e6                                   # "cfg1 ~ 2.0-2.2 MFLOP/graph" at the mean size
    assert 50e3 < roofline.step_bytes([32.2], cfg1) < 60e3         # "54-58 KB/graph"
    fl, by = roofline.ax_kernel_work([2048] * 256, 256, elt=2)     # the dominant launch: U = A.[h|a], bf16
    assert fl == 2.0 * 256 * 2048 ** 2 * 256 and by == 256 * (2048 ** 2 + 2 * 2048 * 256) * 2


def _brute_force(n, D, H, E, L, K):
    """Products of one DiffPool level-0 forward + backward, counted product by product (2mnk flops each)."""
    mm = lambda m, k, nn: 2.0 * m * k * nn
    F, Fa = H * (L - 1) + E, H * (L - 1) + K
    fwd = bwd = 0.0

    def stack(rows, dims, first_level0):
        f = b = 0.0
        for k, (i, o) in enumerate(dims):
            f += mm(rows, rows, i) + mm(rows, i, o)                # U = A.X ; V = U.W
            b += mm(i, rows, o) + mm(rows, o, i)                   # dW = U^T dV ; dU = dV W^T
            if not (k == 0 and first_level0):
                b += mm(rows, rows, i)                             # dX = A^T dU (not for the first layer at level 0)
        return f, b
    dims = lambda i0, last: [(i0, H)] + [(H, H)] * (L - 2) + [(H, last)]
    for f_, b_ in (stack(n, dims(D, E), True), stack(n, dims(D, K), True)):
        fwd, bwd = fwd + f_, bwd + b_
    fwd += mm(n, Fa, K) + mm(K, n, F) + mm(K, n, n) + mm(K, n, K)   # assign_pred, X' = S^T Z, T = S^T A, A' = T S
    bwd += 2 * mm(n, Fa, K) + 2 * mm(n, K, F) + mm(n, K, n) + 2 * mm(n, K, K)   # dWp,dza ; dZ,dS ; A.(S dA'^T) ; T^T dA', S dA'^T
    f_, b_ = stack(K, dims(F, E), False)
    fwd, bwd = fwd + f_, bwd + b_ + sum(mm(K, i, K) for i, _ in dims(F, E))      # + dA' += dU X^T per post-pool layer
    fwd += mm(n, K, n)                                             # link loss: P = S S^T
    bwd += 2 * mm(n, n, K)                                         # (G + G^T) S
    return fwd, bwd


def test_model_matches_brute_force_count():
    for n, D, H, E, L, N, ratio in ((2048, 128, 128, 128, 3, 2048, 0.25), (37, 3, 30, 30, 3, 100, 0.1), (500, 16, 64, 64, 4, 1000, 0.25)):
        cfg = dict(kind='soft', N=N, D=D, H=H, E=E, L=L, ratio=ratio, P=1)
        K = int(N * ratio)
        f, b = roofline.graph_flops(n, cfg)
        bf, bb = _brute_force(n, D, H, E, L, K)
        assert abs(f - bf) <= 1e-9 * bf, (f, bf)
        assert abs(b - bb) <= 1e-9 * bb, (b, bb)
