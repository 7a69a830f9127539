"""CUDA-graph capture of the training step for the launch-bound small-graph regime.

An ENZYMES-sized step (B=20, N=100) is ~170 kernel launches of a few microseconds each: issued one by one
from Python the step is bound by the host (ctypes call + buffer allocation per launch), not by the GPU.
``GraphedTrainStep`` captures train.py:196-210 -- zero_grad -> forward -> loss -> backward -> clip -> Adam --
ONCE per batch shape into a CUDA graph whose inputs are static device buffers, and replays it for every batch
of that shape: all kernels still run every step (nothing is cached or skipped), only the per-launch host work
disappears.  Node counts live in a device int32 buffer (the padding-aware tile skipping reads them on the
device; the link-loss normaliser 1/sum n_b^2 is computed on the device too), so a replay needs no host data
beyond the copies into the static inputs.

Capture rule inherited from PyTorch: no autograd graph built on the default stream with this model may still be
alive when a shape is first captured (its AccumulateGrad nodes would tie the capture to the legacy stream) --
drop earlier outputs / losses first.

The captured kernels are exactly the ones the eager path launches (same C-ABI calls, made once at capture
time); tests/test_gpu_graphed.py checks graph replays against the eager path and the oracle.
"""
import numpy as np
import torch

from . import dp


class GraphedTrainStep:
    def __init__(self, model, lr=1e-3, clip=2.0, linkpred=True, warmup=3):
        self.model, self.clip, self.linkpred, self.warmup = model, clip, linkpred, warmup
        self.params = [p for p in model.parameters() if p.requires_grad]
        # clip + Adam as two kernels of this library over flat buffers; the step counter lives on the device
        # (train.py:173: Adam, lr hard-coded 0.001; train.py:209: clip_grad_norm)
        self.optimizer = dp.FlatAdam(self.params, lr=lr, clip=clip)
        self.grads = self.optimizer.grads.attach(model)   # the backward adds its gradients in one launch
        self.soft = hasattr(model, 'num_pooling')
        self._graphs = {}
        self.replayed_launches = 0           # kernels of this library executed through graph replays

    # ---- one eager step on the static buffers (what gets captured) --------------------------------------
    def _eager(self, st):
        m = self.model
        self.grads.zero()
        if self.soft:
            yp = m(st['x'], st['adj'], st['nb'], assign_x=st['ax'])
            loss = m.loss(yp, st['label'], st['adj'], st['nb']) if (self.linkpred and m.linkpred) else \
                m.loss(yp, st['label'])
        else:
            yp = m(st['x'], st['adj'], st['nb'])
            loss = m.loss(yp, st['label'])
        loss.backward()
        self.optimizer.step()                    # clip_grad_norm folded into the Adam kernel
        return yp, loss

    def _capture(self, key, x, adj, label, assign_x):
        dev = x.device
        # the module stashes outputs of earlier eager calls (assign_tensor, link_loss, ...): they keep those autograd
        # graphs -- and their default-stream AccumulateGrad nodes -- alive, which would poison the capture
        for attr in ('assign_tensor', 'assign_tensors', '_S0', 'link_loss', 'entropy_loss', '_plan'):
            if hasattr(self.model, attr):
                try:
                    delattr(self.model, attr)
                except AttributeError:
                    pass
        st = {'x': torch.empty_like(x), 'adj': torch.empty_like(adj), 'label': torch.empty_like(label),
              'nb': torch.empty(x.shape[0], device=dev, dtype=torch.int32)}
        st['ax'] = st['x'] if assign_x is None else torch.empty_like(assign_x)
        self._fill(st, x, adj, None, label, assign_x, full_nb=adj.shape[1])
        # warm-up on a side stream (allocator + lazy kernel attributes), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        try:        # the flat gradient buffer was created on another stream than the warm-up / capture streams: benign
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        snapshot = [p.detach().clone() for p in self.params]
        # a new batch shape can show up mid-training (the smaller last batch of an epoch: the reference's DataLoader
        # does not drop it): the warm-up steps must leave the accumulated Adam moments and step count untouched too
        opt = self.optimizer
        opt_state = [t.detach().clone() for t in (opt.m, opt.v, opt.step_dev)]
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._eager(st)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # the warm-up steps must not count as training: restore parameters and optimiser state
        with torch.no_grad():
            for p, s in zip(self.params, snapshot):
                p.copy_(s)
            for t, s in zip((opt.m, opt.v, opt.step_dev), opt_state):
                t.copy_(s)
        from ._lib import load
        g = torch.cuda.CUDAGraph()
        n0 = int(load().gp_launch_count())
        with torch.cuda.graph(g):
            yp, loss = self._eager(st)
        st['graph'], st['ypred'], st['loss'] = g, yp, loss
        st['launches'] = int(load().gp_launch_count()) - n0      # this library's kernels inside one replay
        self._graphs[key] = st
        return st

    @staticmethod
    def _fill(st, x, adj, nb, label, assign_x, full_nb=None):
        for k, src in (('x', x), ('adj', adj), ('label', label), ('ax', assign_x)):
            if src is not None and src.data_ptr() != st[k].data_ptr():      # callers may fill the static buffers directly
                st[k].copy_(src, non_blocking=True)
        if nb is None:
            st['nb'].fill_(int(full_nb))
        elif torch.is_tensor(nb):
            st['nb'].copy_(nb.to(torch.int32), non_blocking=True)
        else:
            st['nb'].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(nb, dtype=np.int32))), non_blocking=True)

    def static_inputs(self, x, adj, label, assign_x=None):
        """The static input buffers of the graph for this batch shape (captured on first use): writing a batch
        straight into them (e.g. the host->device copy) and passing them to step() skips the staging copy."""
        if assign_x is x:
            assign_x = None
        key = (tuple(x.shape), tuple(adj.shape), adj.dtype, None if assign_x is None else tuple(assign_x.shape))
        st = self._graphs.get(key) or self._capture(key, x, adj, label, assign_x)
        return st['x'], st['adj'], st['label'], st['ax']

    def step(self, x, adj, batch_num_nodes, label, assign_x=None):
        """One training step.  x / adj / label: CUDA tensors (any batch shape: one graph is captured per shape);
        batch_num_nodes: host array, tensor or None (= every graph has N nodes).
        Returns (ypred, loss): static device tensors, overwritten by the next step of the same shape."""
        if assign_x is x:
            assign_x = None
        key = (tuple(x.shape), tuple(adj.shape), adj.dtype, None if assign_x is None else tuple(assign_x.shape))
        st = self._graphs.get(key)
        if st is None:
            st = self._capture(key, x, adj, label, assign_x)
        self._fill(st, x, adj, batch_num_nodes, label, assign_x, full_nb=adj.shape[1])
        st['graph'].replay()
        self.replayed_launches += st['launches']
        return st['ypred'], st['loss']
