"""GPU: CUDA-graph replay of the training step (graphed.GraphedTrainStep) against the eager path.

The graph holds the same kernels the eager path launches; node counts and the link-loss normaliser live on the
device.  After several steps on DIFFERENT batches of one shape, losses and parameters must agree with an eager
run of train.py:196-210 from the same initial weights (fp32 mode: 2e-5 relative -- the split-K weight-gradient
GEMMs accumulate with atomics, so bit equality is not expected; bf16 mode: 2e-3)."""
import copy

import numpy as np
import pytest
import torch

from helpers import rel_l2, synth_batch

pytestmark = pytest.mark.gpu


def _batches(seed, steps, B, N, D, C, n_min):
    return [synth_batch(seed + i, B, N, D, n_min, N, C, 0.1) for i in range(steps)]


@pytest.mark.parametrize('precision,tol,ptol,N,H', [(0, 2e-5, 2e-5, 48, 16), (1, 2e-3, 1e-1, 128, 32)])
def test_graphed_steps_match_eager(precision, tol, ptol, N, H):
    from graph_pooling_b200 import encoders, graphed
    B, D, C = 6, 5, 3
    torch.manual_seed(1)
    m0 = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25).cuda()
    m0.precision = precision
    me, mg = copy.deepcopy(m0), copy.deepcopy(m0)
    batches = _batches(40, 4, B, N, D, C, 4)
    # eager reference run
    opt = torch.optim.Adam(me.parameters(), lr=1e-3)
    le = []
    for x, adj, nb, label in batches:
        me.zero_grad()
        xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
        yp = me(xc, ac, nb, assign_x=xc)
        loss = me.loss(yp, lc, ac, nb)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(me.parameters(), 2.0)
        opt.step()
        le.append(loss.item())
    # graphed run: one capture, four replays
    gs = graphed.GraphedTrainStep(mg, lr=1e-3, clip=2.0)
    lg = []
    for x, adj, nb, label in batches:
        xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
        yp, loss = gs.step(xc, ac, nb, lc)
        lg.append(loss.item())
    assert len(gs._graphs) == 1
    for a, b in zip(lg, le):
        assert abs(a - b) <= tol * abs(b), (lg, le)
    for (k, p), (_, q) in zip(mg.named_parameters(), me.named_parameters()):
        # bf16 mode: the split-K weight-gradient atomics differ run to run at the 1e-6 level and Adam's
        # g / sqrt(v) turns that into sign flips on noise-level gradients (zero-initialised biases): loose bound
        assert rel_l2(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < ptol, k


def test_graphed_base_encoder_and_no_mask():
    from graph_pooling_b200 import encoders, graphed
    torch.manual_seed(2)
    m0 = encoders.GcnEncoderGraph(7, 20, 24, 3, 3).cuda()
    me, mg = copy.deepcopy(m0), copy.deepcopy(m0)
    x, adj, nb, label = synth_batch(50, 5, 32, 7, 32, 32, 3, 0.15)
    xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
    opt = torch.optim.Adam(me.parameters(), lr=1e-3)
    gs = graphed.GraphedTrainStep(mg)
    for _ in range(3):
        me.zero_grad()
        loss = me.loss(me(xc, ac, None), lc)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(me.parameters(), 2.0)
        opt.step()
        _, lgr = gs.step(xc, ac, None, lc)
        assert abs(lgr.item() - loss.item()) < 2e-5 * abs(loss.item())


def test_second_shape_captured_mid_training_keeps_adam_state(monkeypatch):
    """A new batch shape first seen after some training steps (the smaller last batch of an epoch) is captured
    then: its warm-up steps must not disturb the parameters, the Adam moments or the step count.  Steps
    [full, full, full, small, full] against eager clip_grad_norm + Adam on the same batches.
    Runs on the dense fp32 schedule, whose gradients are bit-reproducible: the packed schedule accumulates parameter
    gradients with atomics, and Adam's g / sqrt(v) turns last-bit differences of near-zero bias gradients into sign
    flips of whole updates (the packed schedule under graph replay is covered by tests/test_gpu_packed.py)."""
    monkeypatch.setenv('GP_NO_PACKED', '1')
    from graph_pooling_b200 import encoders, graphed
    N, H, D, C = 48, 16, 5, 3
    torch.manual_seed(5)
    m0 = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25).cuda()
    me, mg = copy.deepcopy(m0), copy.deepcopy(m0)
    sizes = [6, 6, 6, 3, 6]
    batches = [synth_batch(60 + i, b, N, D, 4, N, C, 0.1) for i, b in enumerate(sizes)]
    opt = torch.optim.Adam(me.parameters(), lr=1e-3)
    gs = graphed.GraphedTrainStep(mg, lr=1e-3, clip=2.0)
    for x, adj, nb, label in batches:
        xc, ac, lc = torch.tensor(x).cuda(), torch.tensor(adj).cuda(), torch.tensor(label).cuda()
        me.zero_grad()
        yp = me(xc, ac, nb, assign_x=xc)
        loss = me.loss(yp, lc, ac, nb)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(me.parameters(), 2.0)
        opt.step()
        del yp
        _, lg = gs.step(xc, ac, nb, lc)
        assert abs(lg.item() - loss.item()) <= 1e-4 * abs(loss.item())
        del loss
    assert len(gs._graphs) == 2 and float(gs.optimizer.step_dev.item()) == float(len(sizes))
    for (k, p), (_, q) in zip(mg.named_parameters(), me.named_parameters()):
        # Adam's g / sqrt(v) amplifies last-bit gradient differences on the near-zero bias gradients (measured 3e-3 on
        # conv_last2.bias); a wiped optimiser state shows up as >= 0.2 on the biases
        assert rel_l2(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 2e-2, k
