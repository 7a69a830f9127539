"""Data-parallel training of the DiffPool encoders: one process per GPU, graphs sharded across ranks, ONE
all-reduce of a flat fp32 gradient buffer per step (NCCL over NVLink on the GPUs; gloo in the CPU tests).
Nothing else crosses GPUs: BatchNorm statistics stay shard-local, exactly as the reference
run on that shard would compute them (encoders.py:1048-1052).

Two loss conventions:
  'shard_mean'   every rank computes the reference loss of its own shard (CE mean over the shard, link loss
                 normalised by the shard's sum of n_b^2, encoders.py:1127,1326-1331); gradients are AVERAGED.
                 G ranks == the mean of G single-GPU reference runs on the shards.
  'global_norm'  the loss is that of the whole batch: CE weighted by shard_size / global_batch and the link
                 loss normalised by the GLOBAL sum of n_b^2 (every rank knows all n_b on the host);
                 gradients are SUMMED.

Parameter gradients live as views of one flat buffer, so the collective needs no gather / scatter copies, and
gradient clipping (train.py:209) uses the norm of the reduced buffer, which every rank holds: no second
collective.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced shard [lo, hi) of `n_items` for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(rank, world, x, adj, batch_num_nodes, label, assign_x=None):
    """Slice one global padded batch (train.py:197-201 tensors) into this rank's shard."""
    lo, hi = shard_bounds(x.shape[0], world, rank)
    nb = None if batch_num_nodes is None else np.asarray(batch_num_nodes)[lo:hi]
    out = dict(x=x[lo:hi], adj=adj[lo:hi], nb=nb, label=label[lo:hi])
    out['assign_x'] = None if assign_x is None else assign_x[lo:hi]
    return out


class FlatGradients:
    """One contiguous fp32 buffer holding every parameter's gradient (p.grad are views into it).

    On CUDA nothing here is an ATen op: zero() is gp_fill_f32, the averaging after the SUM all-reduce is a factor
    (`pending_scale`) that FlatAdam folds into gp_adam_step_f32 (or gp_clip_scale_f32 applies), clip_() is
    gp_sumsq_f32 + gp_clip_scale_f32, and -- once attach()ed to a drop-in encoder -- the backward pass adds all its
    parameter gradients into the buffer with ONE gp_multi_axpy_f32 launch instead of one autograd AccumulateGrad add
    per parameter.  CPU tensors (the gloo tests run the oracle through this class) take the plain torch ops."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev, dt = self.params[0].device, self.params[0].dtype
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        self.cuda = self.flat.is_cuda and dt == torch.float32
        self.pending_scale = 1.0            # factor still to be applied to `flat` (1 / world after a SUM all-reduce)
        self._sumsq = self._ws = None
        self._offset = {}
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self._offset[id(p)] = off
            off += p.numel()

    def attach(self, model):
        """Let `model`'s backward (graph_pooling_b200.encoders) deliver its parameter gradients straight into this
        buffer (accumulating, like autograd would)."""
        if self.cuda:
            model._grad_sink = self
        return self

    def zero(self):
        if self.cuda:
            import ctypes as C
            from ._lib import call
            call('gp_fill_f32', self.flat.data_ptr(), C.c_longlong(self.flat.numel()), C.c_float(0.0),
                 torch.cuda.current_stream().cuda_stream)
        else:
            self.flat.zero_()
        self.pending_scale = 1.0
        off = 0
        for p in self.params:                       # re-attach if someone replaced / dropped p.grad
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def accumulate(self, params, grads):
        """Called by the encoders' backward: adds every gradient whose parameter lives in this buffer with one
        gp_multi_axpy_f32 launch per 64 parameters and returns the list with those entries replaced by None (autograd
        then has nothing left to accumulate for them)."""
        import ctypes as C
        from ._lib import GpAxpyEntry, call
        out, ent = list(grads), []
        base, esz = self.flat.data_ptr(), self.flat.element_size()
        for i, (p, g) in enumerate(zip(params, grads)):
            off = self._offset.get(id(p))
            if g is None or off is None or p.grad is None or p.grad.data_ptr() != base + off * esz:
                continue
            if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == p.numel()):
                continue
            ent.append((g.data_ptr(), base + off * esz, g.numel()))
            out[i] = None
        if ent:
            tab = (GpAxpyEntry * len(ent))(*[GpAxpyEntry(a, b, n) for a, b, n in ent])
            call('gp_multi_axpy_f32', C.cast(tab, C.c_void_p), len(ent), C.c_float(1.0),
                 torch.cuda.current_stream().cuda_stream)
        return out

    def all_reduce(self, group=None, average=True):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                if self.cuda:
                    self.pending_scale = 1.0 / world      # folded into the optimiser / clip kernel
                else:
                    self.flat.div_(world)
        return self.flat

    def apply_pending_scale(self):
        """Materialise the averaging factor in the buffer (callers that read `flat` directly)."""
        if self.pending_scale != 1.0:
            import ctypes as C
            from ._lib import call
            call('gp_clip_scale_f32', self.flat.data_ptr(), C.c_longlong(self.flat.numel()), None, C.c_float(0.0),
                 C.c_float(self.pending_scale), torch.cuda.current_stream().cuda_stream)
            self.pending_scale = 1.0
        return self.flat

    def clip_(self, max_norm):
        """torch.nn.utils.clip_grad_norm_ semantics on the (already reduced) flat buffer; returns the norm."""
        if not self.cuda:
            norm = self.flat.norm(2)
            coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
            self.flat.mul_(coef)
            return norm
        import ctypes as C
        from ._lib import call
        st = torch.cuda.current_stream().cuda_stream
        if self._sumsq is None:
            self._sumsq = torch.zeros(1, device=self.flat.device)
            self._ws = torch.empty(1024, device=self.flat.device)
        n = C.c_longlong(self.flat.numel())
        scale = self.pending_scale
        call('gp_sumsq_f32', self.flat.data_ptr(), n, self._sumsq.data_ptr(), self._ws.data_ptr(), st)
        call('gp_clip_scale_f32', self.flat.data_ptr(), n, self._sumsq.data_ptr(), C.c_float(float(max_norm)),
             C.c_float(scale), st)
        self.pending_scale = 1.0
        return self._sumsq.sqrt() * scale           # diagnostic only (not on the training path)


class FlatAdam:
    """train.py:209-210 (clip_grad_norm + Adam.step) as TWO kernels of this library over flat buffers: parameters are
    re-pointed to views of one flat tensor (state_dict / .parameters() keep working), gradients live in a
    FlatGradients buffer, and gp_adam_step_f32 applies the clip coefficient and torch.optim.Adam's update (lr,
    betas, eps; no weight decay / amsgrad -- train.py:173) in one pass.  The step counter is a device scalar, so the
    whole optimiser step can sit inside a captured CUDA graph.  An all-reduce of `grads.flat` (data parallel) goes
    between backward() and step()."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip=2.0, grads=None):
        from ._lib import call
        self._call = call
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not self.params[0].is_cuda:
            raise ValueError('FlatAdam needs CUDA parameters')
        self.lr, self.betas, self.eps, self.clip = float(lr), betas, float(eps), clip
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                off += k
        self.grads = grads if grads is not None else FlatGradients(self.params)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.step_dev = torch.zeros(1, device=dev)
        self.sumsq = torch.zeros(1, device=dev)
        self._ws = torch.empty(1024, device=dev)

    def step(self):
        import ctypes as C
        st = torch.cuda.current_stream().cuda_stream
        g = self.grads.flat
        n = C.c_longlong(g.numel())
        clip = float(self.clip) if self.clip is not None else 0.0
        if clip > 0:
            self._call('gp_sumsq_f32', g.data_ptr(), n, self.sumsq.data_ptr(), self._ws.data_ptr(), st)
        self._call('gp_adam_step_f32', self.flat_p.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), n,
                   C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                   self.step_dev.data_ptr(), self.sumsq.data_ptr() if clip > 0 else None, C.c_float(clip),
                   C.c_float(self.grads.pending_scale), st)
        self.grads.pending_scale = 1.0


class DataParallelTrainer:
    """Runs train.py:196-210 (zero_grad -> forward -> loss -> backward -> clip -> optimizer step) on this rank's
    shard and synchronises gradients with one all-reduce.  `model` is one of the drop-in encoders (or any module
    with the same forward / loss signatures, e.g. the oracle in the CPU tests)."""

    def __init__(self, model, optimizer=None, clip=2.0, mode='shard_mean', group=None, linkpred=True):
        if mode not in ('shard_mean', 'global_norm'):
            raise ValueError("mode must be 'shard_mean' or 'global_norm'")
        self.model, self.optimizer, self.clip, self.mode, self.group = model, optimizer, clip, mode, group
        self.linkpred = linkpred
        self.grads = FlatGradients(model.parameters()).attach(model)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for p in self.model.parameters():
                dist.broadcast(p.data, src=src, group=self.group)

    def _loss(self, ypred, label, adj, nb, global_nb, global_batch):
        m = self.model
        soft = hasattr(m, 'num_pooling') and getattr(m, 'linkpred', False) and self.linkpred
        if self.mode == 'shard_mean' or self.world == 1:
            return m.loss(ypred, label, adj, nb) if soft else m.loss(ypred, label)
        ce_scale = float(ypred.shape[0]) / float(global_batch)
        if not soft:
            return m.loss(ypred, label) * ce_scale
        g64 = np.asarray(global_nb).astype(np.int64)
        entries_global = int(np.sum(g64 * g64))
        if hasattr(m, 'set_loss_scaling'):             # CUDA encoders: folded into the loss kernels' scalars
            m.set_loss_scaling(ce_scale, entries_global)
            try:
                return m.loss(ypred, label, adj, nb)
            finally:
                m.set_loss_scaling(1.0, None)
        total = m.loss(ypred, label, adj, nb)           # generic module: total = CE + link (both differentiable)
        l64 = np.asarray(nb).astype(np.int64)
        link = m.link_loss
        return (total - link) * ce_scale + link * (float(np.sum(l64 * l64)) / float(entries_global))

    def step(self, x, adj, batch_num_nodes, label, assign_x=None, global_num_nodes=None, global_batch=None):
        """One training step on this rank's shard.  global_num_nodes / global_batch: every rank's n_b (host) and
        the global batch size; needed by mode='global_norm' only.  Returns (ypred, loss) of the shard."""
        self.grads.zero()
        kw = {} if assign_x is None else {'assign_x': assign_x}
        ypred = self.model(x, adj, batch_num_nodes, **kw)
        if self.mode == 'global_norm' and self.world > 1 and (global_num_nodes is None or global_batch is None):
            raise ValueError("mode='global_norm' needs global_num_nodes and global_batch")
        loss = self._loss(ypred, label, adj, batch_num_nodes, global_num_nodes, global_batch)
        loss.backward()
        self.grads.all_reduce(self.group, average=(self.mode == 'shard_mean'))
        if isinstance(self.optimizer, FlatAdam) and self.optimizer.grads is self.grads:
            self.optimizer.clip = self.clip            # averaging factor + clip folded into the Adam kernel
            self.optimizer.step()
            return ypred, loss
        if self.clip is not None:
            self.grads.clip_(self.clip)
        else:
            self.grads.apply_pending_scale() if self.grads.cuda else None
        if self.optimizer is not None:
            self.optimizer.step()
        return ypred, loss
