#!/usr/bin/env python
"""bench.py -- DiffPool train-step throughput (graphs/s) on B200, with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is train.py:196-210 on one synthetic batch: zero_grad -> forward -> loss (CE + link
prediction) -> backward -> clip_grad_norm(2.0) -> Adam(lr=1e-3).  `value` times it with the
batch resident in HBM; `e2e` times the same step through the drop-in encoders with HOST (pinned)
fp32 buffers, host->device copies and the loss read-back inside the timed region.
N > 1 (torchrun): one process per GPU, each with its own batch of the same shape (weak scaling),
one NCCL all-reduce of the flat gradient per step; time = max over ranks.

--impl reference: the oracle restatement of the reference's PyTorch code (oracle/, the reference
itself does not construct -- SURVEY.md 0-3) on the host CPU with all threads, on a bounded sample
of the same workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'diffpool_train_graphs_per_sec'
UNIT = 'graphs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg4_diffpool_256x2048')
    ap.add_argument('--batch', type=int, default=None, help='override graphs per GPU per step')
    ap.add_argument('--cpu-sample', type=int, default=None, help='graphs per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--graph', default='auto', choices=['auto', 'on', 'off'],
                    help='replay the step from a CUDA graph (graphed.GraphedTrainStep); auto: on when the batch '
                         'adjacency is <= 256 MB (launch-bound steps), off for the multi-GB batches')
    ap.add_argument('--precision', default='auto', choices=['auto', 'f32', 'bf16'],
                    help='auto: bf16 tensor-core path for N >= 512 (BASELINE configs[3]), fp32 FFMA otherwise')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.06)                                  # let the sample that covers the end of the region arrive
        self.proc.terminate()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

        def collect(lo, hi):
            sm, smax, reasons = [], None, set()
            for t, line in self.rows:
                if t < lo or t > hi:
                    continue
                f = [x.strip() for x in line.split(',')]
                try:
                    sm.append(float(f[0]))
                    smax = float(f[1])
                except Exception:
                    continue
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            return sm, smax, reasons

        sm, smax, reasons = collect(t0, t1 + 0.15)
        window = 'timed region'
        if not sm:       # a region shorter than the sampling period: fall back to warm-up + region (same load)
            sm, smax, reasons = collect(0.0, t1 + 0.15)
            window = 'warm-up + timed region (region shorter than the sampling period)'
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm), 'window': window}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d['bf16_tflops_sustained'], d['bf16_tflops'], 'measured'
    return 6650.0, 1400.0, 1590.0, 'fallback'


# ------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """Oracle (port of the reference) on the host CPU; rank 0 only."""
    if rank != 0:
        return
    from graph_pooling_b200 import synth
    from oracle import diffpool_oracle as orc
    cfg = dict(synth.WORKLOADS[args.workload])
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = args.cpu_sample or default_cpu_sample(cfg)
    value, ms, desc = time_cpu_oracle(args.workload, sample, args.steps, args.warmup, args.seed)
    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
           'config': {'workload': args.workload, 'graphs_per_step': sample, 'sample': desc},
           'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
           'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out), flush=True)


def default_cpu_sample(cfg):
    # ~10-30 s of CPU work in total: scale the per-step sample with the per-graph cost; never fewer than 8 graphs
    # (BatchNorm couples the graphs of a batch: a 2-graph sample is not the same computation per graph)
    cost = cfg['N'] * cfg['N'] * (cfg['H'] * 6 + int(cfg['N'] * cfg['ratio']) * 6)
    return int(min(cfg['B'], max(8, 1.6e11 / max(cost, 1))))


def time_oracle(workload, sample, steps, warmup, seed, device='cpu', with_optimizer=True):
    """The oracle restatement of the reference (plain torch ops + autograd) on `device`: 'cpu' = the reference's CPU
    path on the host cores; a CUDA device = the "ATen-on-B200" comparator (torch eager: cuBLAS / ATen kernels)."""
    from graph_pooling_b200 import synth
    from oracle import diffpool_oracle as orc
    batch = synth.make_batch(workload, seed=seed, device=device, B=sample)
    cfg = batch['cfg']
    torch.manual_seed(seed)
    model = synth.build_model(orc, cfg).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3) if with_optimizer else None
    x, adj, nb, label = batch['x'], batch['adj'], batch['nb'], batch['label']
    soft = cfg['kind'] == 'soft'
    cuda = torch.device(device).type == 'cuda'
    ts = []
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, loss = orc.train_step(model, x, adj, label, nb, assign_x=x if soft else None, optimizer=opt)
        if cuda:
            float(loss.item())
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts[warmup:]))
    desc = ('%d graphs/step of %s (same shapes), %d warm-up + %d timed steps, %s, torch %s %s fp32%s'
            % (sample, workload, warmup, steps, 'zero_grad+fwd+loss+bwd+clip+Adam' if with_optimizer else
               'zero_grad+fwd+loss+bwd only', torch.__version__,
               'eager on the GPU (cuBLAS/ATen kernels)' if cuda else 'CPU',
               '' if cuda else ', %d threads' % torch.get_num_threads()))
    return sample / t, t * 1e3, desc


def time_cpu_oracle(workload, sample, steps, warmup, seed):
    return time_oracle(workload, sample, steps, warmup, seed, 'cpu', True)


def enzymes_regime_point(dev, hbm, B=4096, reps=20):
    """The small-graph half of BASELINE.json's metric inside the driver-run record: cfg1 (ENZYMES-like) shapes at
    B = 4096 graphs, fp32 schedule, whole train step replayed from a CUDA graph, batch resident; algorithmic bytes
    (roofline.step_bytes, fp32) / time against the measured HBM peak."""
    from graph_pooling_b200 import encoders, graphed, roofline, synth
    batch = synth.make_batch('cfg1_enzymes_like', seed=0, device=dev, B=B)
    cfg = batch['cfg']
    torch.manual_seed(0)
    model = synth.build_model(encoders, cfg).to(dev)
    model.precision = 0
    gs = graphed.GraphedTrainStep(model, lr=1e-3, clip=2.0)
    nbd = torch.from_numpy(np.ascontiguousarray(batch['nb'].astype(np.int32))).to(dev)
    x, adj, label = batch['x'], batch['adj'], batch['label']
    for _ in range(3):
        gs.step(x, adj, nbd, label)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gs.step(x, adj, nbd, label)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = roofline.step_bytes(batch['nb'], cfg) / 1e9
    st = next(iter(gs._graphs.values()))
    return {'workload': 'cfg1_enzymes_like', 'graphs_per_step': B, 'precision': 'f32', 'cuda_graph': True,
            'ms_per_step': ms, 'graphs_per_s': B / (ms * 1e-3), 'launches_per_step': st['launches'],
            'algorithmic_gb_per_step': gb, 'achieved_gbs': gb / (ms * 1e-3), 'peak_gbs': hbm,
            'frac_of_hbm_peak': gb / (ms * 1e-3) / hbm,
            'note': 'bytes counted at fp32 over the real n_b x n_b blocks and real rows (SURVEY 8(d)); inputs resident, '
                    'smaller than L2 per graph but %.0f MB per batch' % (adj.numel() * 4 / 1e6)}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        reference_arm(args, rank, world)
        return

    from graph_pooling_b200 import _lib, encoders, roofline, synth
    from graph_pooling_b200 import engine as E
    assert torch.cuda.is_available(), 'bench.py (impl=ours) needs a CUDA device: there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    batch = synth.make_batch(args.workload, seed=args.seed + rank, device=dev, B=args.batch)
    cfg = batch['cfg']
    soft = cfg['kind'] == 'soft'
    torch.manual_seed(args.seed)
    model = synth.build_model(encoders, cfg).to(dev)
    prec = args.precision if args.precision != 'auto' else ('bf16' if cfg['N'] >= 512 else 'f32')
    model.precision = 1 if prec == 'bf16' else 0
    params = [p for p in model.parameters()]
    x, adj, nb, label = batch['x'], batch['adj'], batch['nb'], batch['label']
    B = x.shape[0]

    # gradients live in ONE flat buffer (dp.FlatGradients): the all-reduce and the clip need no gather copies
    from graph_pooling_b200 import dp
    use_graph = (args.graph == 'on' or (args.graph == 'auto' and adj.numel() * 4 <= 256e6)) and world == 1
    gstep = None
    if not use_graph:
        opt = dp.FlatAdam(params, lr=1e-3, clip=2.0)     # clip_grad_norm + Adam: two kernels of this library
        flat = opt.grads.attach(model)                   # the backward adds all parameter gradients in ONE launch
    if use_graph:
        from graph_pooling_b200 import graphed
        gstep = graphed.GraphedTrainStep(model, lr=1e-3, clip=2.0)
        flat, opt = gstep.grads, gstep.optimizer
        nb_dev = torch.from_numpy(np.ascontiguousarray(nb.astype(np.int32))).to(dev)

    def step(xd, ad, ld):
        if gstep is not None:                       # all kernels replayed from one CUDA graph, node counts on device
            return gstep.step(xd, ad, nb_dev, ld)[1]
        flat.zero()
        yp = model(xd, ad, nb, assign_x=xd) if soft else model(xd, ad, nb)
        loss = model.loss(yp, ld, ad, nb) if soft else model.loss(yp, ld)
        loss.backward()
        flat.all_reduce(average=True)                # one NCCL all-reduce per step (no-op at world == 1)
        opt.step()                                   # train.py:209-210: clip_grad_norm(2.0) folded into Adam
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None      # started early: already streaming when timing begins
    xs_, adjs_, labels_ = x, adj, label
    if gstep is not None and os.environ.get('GP_BENCH_STATIC', '1') != '0':
        # the resident batch lives in the captured graph's own input buffers: no staging copy inside the step
        xs_, adjs_, labels_, _ = gstep.static_inputs(x, adj, label)
        xs_.copy_(x); adjs_.copy_(adj); labels_.copy_(label)
    for _ in range(args.warmup):
        step(xs_, adjs_, labels_)
    lib.gp_launch_count_reset()
    if gstep is not None:
        gstep.replayed_launches = 0
    t0 = time.time()
    if os.environ.get('GP_PROFILE'):            # ncu --profile-from-start off: capture the timed steps only
        torch.cuda.cudart().cudaProfilerStart()
    ms = timed(lambda: step(xs_, adjs_, labels_), args.steps)
    if os.environ.get('GP_PROFILE'):
        torch.cuda.cudart().cudaProfilerStop()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = int(lib.gp_launch_count()) + (gstep.replayed_launches if gstep is not None else 0)
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end: host pinned buffers -> H2D -> step -> loss read-back ---------------------------
    # Double-buffered: the copy of step i+1's inputs runs on a copy stream while step i computes (what a
    # DataLoader with pin_memory + non_blocking .cuda() gives train.py:197-201); every step's inputs cross PCIe
    # and every step's loss is read back inside the timed region.
    e2e = None
    e2e_u8 = None
    e2e_edges = None
    if not args.no_e2e:
        def run_e2e(adj_host_dtype):
            hx, hl = x.cpu().pin_memory(), label.cpu().pin_memory()
            ha = (adj.to(adj_host_dtype) if adj_host_dtype != adj.dtype else adj).cpu().pin_memory()
            bufs = [(torch.empty_like(x), torch.empty(adj.shape, device=dev, dtype=adj_host_dtype),
                     torch.empty_like(label)) for _ in range(2)]
            copy_stream = torch.cuda.Stream(device=dev)
            ready = [torch.cuda.Event(), torch.cuda.Event()]
            freed = [torch.cuda.Event(), torch.cuda.Event()]
            state = {'i': 0}

            def issue(slot):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[slot])          # the step that last used this slot is done
                    dx, da, dl = bufs[slot]
                    dx.copy_(hx, non_blocking=True)
                    da.copy_(ha, non_blocking=True)
                    dl.copy_(hl, non_blocking=True)
                    ready[slot].record(copy_stream)

            for sl in range(2):
                freed[sl].record(torch.cuda.current_stream())
            issue(0)

            def e2e_step():
                slot = state['i'] & 1
                state['i'] += 1
                issue(slot ^ 1)                                  # prefetch the next step's inputs
                torch.cuda.current_stream().wait_event(ready[slot])
                dx, da, dl = bufs[slot]
                loss = step(dx, da, dl)
                freed[slot].record(torch.cuda.current_stream())
                return float(loss.item())                        # D2H read-back of the step's result

            for _ in range(min(args.warmup, 3)):
                e2e_step()
            ems = timed(e2e_step, args.steps) / args.steps
            torch.cuda.synchronize()
            h2d = hx.numel() * 4 + ha.numel() * ha.element_size() + hl.numel() * 8 + nb.nbytes
            return {'value': world * B / (ems * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ems}

        def run_e2e_hybrid():
            """feed.HostAdjacencyFeed: same fp32 host buffers as run_e2e; the host cores bit-pack part of the batch
            while the rest crosses PCIe as fp32 (pack of step i+2, H2D of step i+1, compute of step i overlap)."""
            from graph_pooling_b200 import feed
            Nn = cfg['N']
            threads = max(1, (os.cpu_count() or 1) // max(world, 1))
            hx, hl, ha = x.cpu().pin_memory(), label.cpu().pin_memory(), adj.cpu().pin_memory()
            nbd_ = torch.from_numpy(np.ascontiguousarray(nb.astype(np.int32))).to(dev)
            t_pack, t_h2d = feed.HostAdjacencyFeed.measure_rates(ha, dev, threads)   # all ranks at once: same contention
            f0 = t_h2d / (t_pack + t_h2d)

            def make(Bp):
                fd = feed.HostAdjacencyFeed(B, Nn, dev, packed_graphs=Bp, threads=threads)
                st = {'fd': fd, 'i': 0, 'xd': [torch.empty_like(x) for _ in range(2)],
                      'ld': [torch.empty_like(label) for _ in range(2)]}
                fd.submit(ha, 0)
                fd.copy(0, [(st['xd'][0], hx), (st['ld'][0], hl)])
                fd.submit(ha, 1)
                return st

            def hstep(st):
                fd, i = st['fd'], st['i']
                st['i'] += 1
                cur, nxt = i & 1, (i + 1) & 1
                fd.copy(nxt, [(st['xd'][nxt], hx), (st['ld'][nxt], hl)])     # H2D of step i+1
                fd.submit(ha, cur)                                              # host pack of step i+2
                pa = fd.prepared(cur, nbd_)
                loss = step(st['xd'][cur], pa, st['ld'][cur])
                fd.consumed(cur)                                                # next copy into this slot waits for the step
                return float(loss.item())                                       # compute of step i + D2H read-back

            best = None
            for f in sorted({min(1.0, max(0.0, f0 + df)) for df in (-0.12, 0.0, 0.12, 0.24)}):
                st = make(int(round(f * B)))
                hstep(st)
                torch.cuda.synchronize(); t0_ = time.perf_counter()
                for _ in range(2):
                    hstep(st)
                torch.cuda.synchronize(); dt = (time.perf_counter() - t0_) / 2
                st['fd'].close()
                if best is None or dt < best[0]:
                    best = (dt, st['fd'].Bp)
                del st
            Bp = best[1]
            st = make(Bp)
            for _ in range(min(args.warmup, 3)):
                hstep(st)
            ems = timed(lambda: hstep(st), args.steps) / args.steps
            torch.cuda.synchronize()
            st['fd'].close()
            ldb = (Nn + 7) // 8
            h2d = hx.numel() * 4 + hl.numel() * 8 + nb.nbytes + Bp * Nn * ldb + (B - Bp) * Nn * Nn * 4
            return {'value': world * B / (ems * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ems,
                    'strategy': 'feed.HostAdjacencyFeed: hybrid host bit-pack + fp32 copy',
                    'packed_graphs_per_step': Bp, 'host_pack_threads': threads,
                    'host_pack_gbs': adj.numel() * 4 / t_pack / 1e9, 'h2d_gbs': adj.numel() * 4 / t_h2d / 1e9,
                    'note': 'fp32 adjacency in pinned host memory every step (the reference feed contract); the plugin '
                            'bit-packs %d of %d graphs on %d host threads (exact for {0,1}) while the other graphs '
                            'cross PCIe as fp32; pack of step i+2, H2D of step i+1 and compute of step i overlap'
                            % (Bp, B, threads)}

        def run_e2e_edges():
            """feed.EdgeListFeed (the plugin's own feed API): every step the host hands over the batch's
            edge lists (pinned int32) and features; gp_adj_from_edges builds the bf16 operand on the device."""
            from graph_pooling_b200 import feed
            Nn = cfg['N']
            au = torch.triu(adj, diagonal=1)
            eptr_l, chunks = [0], []
            for b in range(B):                                   # per graph: bounded temporaries
                nz = au[b].nonzero().to(torch.int32)
                chunks.append(nz)
                eptr_l.append(eptr_l[-1] + int(nz.shape[0]))
            del au
            # compact encodings the feed API accepts (done ONCE, outside the timed region -- a property of how the dataset
            # is stored, like the edge lists themselves): 16-bit node ids when N allows, features in bf16 when the model
            # computes in bf16 (the tensor-core schedule reads nothing else)
            e_dt = torch.int16 if Nn <= 32768 else torch.int32
            x_bf16 = prec == 'bf16' and x.shape[2] % 8 == 0
            he = torch.cat(chunks).to(e_dt).cpu().contiguous().pin_memory()
            hp = torch.tensor(eptr_l, dtype=torch.int32).pin_memory()
            maxdeg = int(max(np.diff(np.asarray(eptr_l)))) if B else 1
            xsrc = x.to(torch.bfloat16) if x_bf16 else x
            hx, hl = xsrc.cpu().pin_memory(), label.cpu().pin_memory()
            fd = feed.EdgeListFeed(B, Nn, dev, max_edges=max(int(he.shape[0]), 1), edge_dtype=e_dt)
            xd = [torch.empty_like(xsrc) for _ in range(2)]
            ld_ = [torch.empty_like(label) for _ in range(2)]
            st = {'i': 0}
            fd.copy(0, he, hp, maxdeg, [(xd[0], hx), (ld_[0], hl)])

            # the step's loss is read back EVERY step, asynchronously: D2H into a pinned slot + event on the compute
            # stream, consumed one step later (how a training loop logs losses without draining the GPU; the reference
            # itself only accumulates the loss tensor, train.py:211).  The last one is collected inside the timed region.
            hloss = [torch.empty(1).pin_memory() for _ in range(2)]
            levt = [torch.cuda.Event(), torch.cuda.Event()]
            seen = {'loss': None, 'pending': None}

            def collect():
                if seen['pending'] is not None:
                    levt[seen['pending']].synchronize()
                    seen['loss'] = float(hloss[seen['pending']][0])
                    seen['pending'] = None

            def estep():
                i = st['i']
                st['i'] += 1
                cur, nxt = i & 1, (i + 1) & 1
                fd.copy(nxt, he, hp, maxdeg, [(xd[nxt], hx), (ld_[nxt], hl)])     # H2D of step i+1
                pa = fd.prepared(cur)
                loss = step(xd[cur], pa, ld_[cur])
                fd.consumed(cur)
                collect()                                                          # loss of step i-1
                hloss[cur].copy_(loss.detach().reshape(1), non_blocking=True)
                levt[cur].record(torch.cuda.current_stream())
                seen['pending'] = cur

            def erun():
                estep()

            for _ in range(min(args.warmup, 3)):
                estep()
            collect()

            def timed_all():
                for _ in range(args.steps):
                    estep()
                collect()

            ems = timed(timed_all, 1) / args.steps
            torch.cuda.synchronize()
            assert seen['loss'] is not None and np.isfinite(seen['loss'])
            h2d = hx.numel() * hx.element_size() + hl.numel() * 8 + nb.nbytes + he.numel() * he.element_size() + hp.numel() * 4
            return {'value': world * B / (ems * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ems, 'edges_per_step': int(he.shape[0]),
                    'edge_id_bytes': int(he.element_size()), 'feature_dtype': str(hx.dtype).replace('torch.', ''),
                    'strategy': 'feed.EdgeListFeed: pinned edge lists + features -> H2D -> gp_adj_from_edges',
                    'loss_readback': 'every step, asynchronous (pinned D2H + event, consumed one step later)',
                    'note': 'the plugin\'s own feed API (SURVEY 8(f) N2): the adjacency crosses PCIe as 8 bytes per '
                            'undirected edge instead of 4 N^2 bytes per graph; same step'}

        e2e = run_e2e(torch.float32)
        e2e['note'] = ('dense fp32 adjacency from pinned host memory every step (the reference feed contract, '
                       'train.py:197-201); H2D of step i+1 overlaps step i on a copy stream')
        e2e['strategy'] = 'direct fp32 copy'
        if prec == 'bf16' and gstep is None:
            # SURVEY 8(f) N2 preview: the same step fed a uint8 {0,1} adjacency (a quarter of the PCIe bytes);
            # NOT the headline e2e -- the reference's train.py feeds fp32.
            e2e_u8 = run_e2e(torch.uint8)
            e2e_u8['note'] = 'uint8 adjacency feed (compact-feed extension, not the reference contract)'
            # Same fp32 host buffers as `e2e`, smarter plugin: the step is PCIe-bound (4.3 GB of {0,1} floats), so the
            # host cores bit-pack a fraction of the graphs (32x smaller, exact) WHILE the rest crosses PCIe as fp32;
            # the fraction comes from the two rates measured here.  Every step packs, copies and computes its own batch.
            try:
                hyb = run_e2e_hybrid()
            except Exception as ex:            # e.g. entries outside {0,1}: the direct copy stays the answer
                hyb = {'error': str(ex)[:200]}
            try:
                e2e_edges = run_e2e_edges()
            except Exception as ex:
                e2e_edges = {'error': str(ex)[:200]}
            if 'value' in hyb and hyb['value'] > e2e['value']:
                hyb['direct_fp32_copy'] = {k: e2e[k] for k in ('value', 'ms_per_step', 'h2d_bytes_per_step')}
                e2e = hyb
            else:
                e2e['hybrid_host_pack'] = hyb

    # ---- data-parallel correctness on the NCCL path (outside every timed region) ---------------------------
    dp_check = None
    if world > 1:
        # (1) one step's reduced gradient == mean of the shard gradients gathered from every rank
        flat.zero()
        yp = model(x, adj, nb, assign_x=x) if soft else model(x, adj, nb)
        (model.loss(yp, label, adj, nb) if soft else model.loss(yp, label)).backward()
        mine = flat.flat.clone()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        mean = torch.stack(parts).double().mean(0)
        flat.all_reduce(average=True)
        flat.apply_pending_scale()
        err = float((flat.flat.double() - mean).norm() / mean.norm().clamp_min(1e-30))
        # (2) the replicas are bit-identical after all the steps above: MAX - MIN over ranks of every parameter == 0
        pmax, pmin = opt.flat_p.clone(), opt.flat_p.clone()
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        spread = float((pmax - pmin).abs().max())
        dp_check = {'reduced_grad_vs_mean_of_shard_grads_rel_l2': err, 'replica_param_max_minus_min': spread,
                    'replicas_bit_identical': spread == 0.0, 'ranks': world}
        assert err < 1e-5 and spread == 0.0, dp_check
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    def timed_local(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    hbm, tf_sus, tf_burst, src = peaks()
    # ---- per-call table of ONE extra (untimed) step: every C-ABI call bracketed by CUDA events -----------------
    from graph_pooling_b200 import profile as gprof
    FFMA_TF = 148 * 128 * 2 * 1.965e9 / 1e12            # fp32 FMA peak of the SIMT pipe (fp32 schedule's compute roof)
    kernels_tab, prof_ms = None, None
    try:
        with gprof.CallProfiler() as cp:
            if gstep is not None:
                gstep._eager(next(iter(gstep._graphs.values())))
            else:
                step(x, adj, label)
        rows_, prof_ms = cp.table(hbm, tf_burst if prec == 'bf16' else FFMA_TF,
                                  min_share=float(os.environ.get('GP_BENCH_MIN_SHARE', 0.03)))
        kernels_tab = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()
                        if k in ('entry', 'shape', 'launches', 'ms', 'ms_per_launch', 'share', 'bound', 'achieved',
                                 'peak', 'unit', 'frac', 'flops', 'bytes', 'counted_at')} for r in rows_]
        for r in kernels_tab:
            if r.get('bound') == 'tensor' and prec != 'bf16':
                r['bound'] = 'ffma'
    except Exception as ex:                    # the table is diagnostic: it must never cost the bench line
        kernels_tab = [{'error': str(ex)[:200]}]

    # ---- the dominant kernel timed alone: the batched A.X contraction -------------------------
    H = cfg['H']
    xin = torch.randn(B, cfg['N'], H, device=dev)
    u = torch.empty(B, cfg['N'], H, device=dev)
    nbd, _ = E.prep_nb(nb, cfg['N'], dev)
    N = cfg['N']

    if prec == 'bf16':
        from graph_pooling_b200 import engine_tc as T
        wsb = E.Workspace(dev)
        adjb = T.cvt(wsb, adj.data_ptr(), N, B * N, N, B=B)
        xinb = T.cvt(wsb, xin.data_ptr(), H, B * N, H, B=B)
        ub = T.bfbuf(wsb, B, N, H)

        def ax():
            T.tcgemm(adjb, T.KM, xinb, T.MN, N, H, N, B, Cb=ub, lim=nbd.data_ptr(), lim_m=1, lim_k=1)
    else:
        def ax():
            E.bgemm(adj.data_ptr(), xin.data_ptr(), u.data_ptr(), N, H, N, B, (N * N, N, 1), (N * H, H, 1),
                    (N * H, H, 1), lim=nbd.data_ptr(), lim_m=1, lim_k=1)
    # ENZYMES-sized graphs (fp32 schedule): the step does not launch the batched GEMM for its GraphConvs but the
    # per-graph fused layer kernel (U = A.X, V = U.W + b, normalize in one CTA per graph) -- that is the launch timed
    # as the dominant one below
    small_fused = prec != 'bf16' and N <= 128 and H <= 128 and not os.environ.get('GP_NO_SMALL_GCN')
    if small_fused:
        from graph_pooling_b200._lib import call as _call
        wsm = torch.randn(H, H, device=dev) * 0.1
        bsm = torch.zeros(H, device=dev)
        ysm, rsm = torch.empty(B, N, H, device=dev), torch.empty(B, N, device=dev)

        def ax_small():
            _call('gp_graphconv_fwd', xin.data_ptr(), H, adj.data_ptr(), wsm.data_ptr(), bsm.data_ptr(),
                  nbd.data_ptr(), B, N, H, H, 0, 1, u.data_ptr(), ysm.data_ptr(), H, rsm.data_ptr(), 0,
                  torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        ax()
    reps = 10
    kms = timed_local(lambda: ax(), reps) / reps
    # the same contraction with the padding-aware schedule switched off (every tile of the padded N x N computed):
    # identical result (the padding is zero), reported to show what the tile skip saves on ragged batches
    if prec == 'bf16':
        def ax_dense():
            T.tcgemm(adjb, T.KM, xinb, T.MN, N, H, N, B, Cb=ub)
    else:
        def ax_dense():
            E.bgemm(adj.data_ptr(), xin.data_ptr(), u.data_ptr(), N, H, N, B, (N * N, N, 1), (N * H, H, 1),
                    (N * H, H, 1))
    ax_dense()
    kms_dense = timed_local(ax_dense, reps) / reps
    # The dominant launch shape of the step.  bf16 DiffPool: embedding and assignment GCN run in lock-step, so the
    # N x N contraction is U = A.[h | a] at 2H columns (4 launches / ~11 % of the cfg4 step, the largest share of one
    # shape; profiles/r1u_launches_cfg4_bf16.md) -- timed live here.  Otherwise: the H-column U = A.X.
    dual = prec == 'bf16' and soft and H % 8 == 0 and H <= 256
    cols = 2 * H if dual else H
    if dual:
        x2 = T.bfbuf(wsb, B, N, cols)
        x2.t.normal_()
        u2 = T.bfbuf(wsb, B, N, cols)

        def ax2():
            T.tcgemm(adjb, T.KM, x2, T.MN, N, cols, N, B, Cb=u2, lim=nbd.data_ptr(), lim_m=1, lim_k=1)
        for _ in range(3):
            ax2()
        kms_dom = timed_local(ax2, reps) / reps
        del x2, u2
    elif small_fused:
        for _ in range(3):
            ax_small()
        kms_dom = timed_local(ax_small, reps) / reps
    else:
        kms_dom = kms
    kfl, kby = roofline.ax_kernel_work(nb, cols, elt=2 if prec == 'bf16' else 4)
    if small_fused:                     # + V = U.W (2 n H^2 flops) and the Y write (n H floats) of the fused layer
        nbf = np.asarray(nb, dtype=np.float64)
        kfl += float(np.sum(2 * nbf * H * H))
        kby += float(np.sum(nbf * H * 4))
    ai = kfl / kby
    ridge = tf_sus * 1e12 / (hbm * 1e9)
    # which roof binds THIS launch: the larger of its HBM time and its tensor time at the measured peaks
    t_hbm, t_tc = kby / (hbm * 1e9), kfl / (tf_burst * 1e12)
    if prec == 'bf16' and t_tc > t_hbm:
        roof = {'bound': 'tensor', 'achieved': kfl / (kms_dom * 1e-3) / 1e12, 'peak': tf_burst, 'unit': 'TFLOP/s'}
    else:
        roof = {'bound': 'hbm', 'achieved': kby / (kms_dom * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s'}
    roof['frac'] = roof['achieved'] / roof['peak']
    traffic = None
    # ncu --set full captures of this shape (profiles/): valid only for the kernel source they were taken on --
    # the file records the sha256 of csrc/gemm_tc2.cu at capture time; a different source => traffic is null
    tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    import hashlib
    ksrc = os.path.join(ROOT, 'graph_pooling_b200', 'csrc', 'gemm_tc2.cu')
    ksha = hashlib.sha256(open(ksrc, 'rb').read()).hexdigest()[:16] if os.path.exists(ksrc) else None
    same_rev = tj.get('kernel_source_sha256_16') == ksha
    tkey = 'ax2_gemm' if dual else 'ax_gemm'
    if prec == 'bf16' and args.workload == 'cfg4_diffpool_256x2048' and B == 256 and tkey in tj and same_rev:
        traffic = tj[tkey]['dram_bytes_per_launch']               # ncu --set full, same shape (profiles/)
        roof['ncu_tensor_pipe_active_pct'] = tj[tkey].get('tensor_pipe_active_pct')
    roof['traffic'] = traffic
    roof['traffic_source'] = (os.path.basename(tpath) + (' (same kernel source)' if same_rev else
                                                         ' is from another revision of the kernel: not reported'))
    roof['algorithmic_bytes_per_launch'] = kby
    roof['algorithmic_flops_per_launch'] = kfl
    roof['arithmetic_intensity_flop_per_byte'] = ai
    roof['kernel'] = '%s (U = A.%s, N=%d, %d columns, batch=%d)' % (
        ('gp::v2::tc_gemm2_kernel<%s> tcgen05+TMA persistent' % ('256,6,0,8,2 (cta_group::2 CTA pair, UMMA M=256)' if cols > 128 else '128,6,0,8'))
        if prec == 'bf16' else ('gp::gconv_small_fwd_kernel: fused per-graph GraphConv layer, V = (A.X).W + b, normalize;'
                                if small_fused else 'gp::bgemm_kernel FFMA'), '[h|a]' if dual else 'X', N, cols, B)
    roof['tflops'] = kfl / (kms_dom * 1e-3) / 1e12
    roof['frac_of_tensor_peak'] = roof['tflops'] / tf_burst if prec == 'bf16' else None
    roof['frac_of_hbm_peak'] = kby / (kms_dom * 1e-3) / 1e9 / hbm
    roof['h_column_ax'] = {'ms_per_launch': kms, 'gbs_algorithmic':
                           roofline.ax_kernel_work(nb, H, elt=2 if prec == 'bf16' else 4)[1] / (kms * 1e-3) / 1e9}
    roof['ms_per_launch'] = kms_dom
    roof['h_column_ax']['ms_per_launch_without_tile_skip'] = kms_dense
    nbf_ = np.asarray(nb, dtype=np.float64)
    roof['occupancy_sum_nb2_over_B_N2'] = float(np.sum(nbf_ * nbf_) / (len(nbf_) * float(N) * N))
    roof['peak_source'] = src + ' (MEASURED_PEAKS.json; kernel timed alone -> burst figures)'
    fwd_fl, bwd_fl = roofline.step_flops(nb, cfg)
    roof['step_algorithmic_tflop'] = (fwd_fl + bwd_fl) / 1e12
    roof['step_tflops'] = (fwd_fl + bwd_fl) / (ms_per_step * 1e-3) / 1e12
    roof['step_frac_of_sustained_bf16'] = roof['step_tflops'] / tf_sus
    roof['step_algorithmic_gb'] = roofline.step_bytes(nb, cfg) / 1e9
    roof['step_gbs'] = roof['step_algorithmic_gb'] / (ms_per_step * 1e-3)
    # the large-N tensor-bound contraction of the pooling step, T = S^T A (encoders.py:1279), timed live too
    if prec == 'bf16' and soft:
        K0 = int(N * cfg['ratio'])
        sop = T.bfbuf(wsb, B, N, K0)                       # row stride padded to 8 elements (TMA 16-byte rule)
        sprob = sop.t
        sprob.copy_(torch.rand(B, N, sprob.shape[2], device=dev))
        sprob[:, :, K0:] = 0
        tbuf = T.bfbuf(wsb, B, K0, N)

        def tsa():
            T.tcgemm(sop, T.MN, adjb, T.MN, K0, N, N, B, Cb=tbuf, lim=nbd.data_ptr(), lim_k=1, lim_n=1)
        for _ in range(3):
            tsa()
        tms = timed_local(tsa, reps) / reps
        nbf = np.asarray(nb, dtype=np.float64)
        tfl = float(np.sum(2.0 * K0 * nbf * nbf))
        roof['tensor_contraction'] = {
            'kernel': 'gp::v2::tc_gemm2_kernel<256,6,0,8,2> cta_group::2 CTA pair (T = S^T.A, K=%d, N=%d, batch=%d)' % (K0, N, B),
            'bound': 'tensor', 'achieved': tfl / (tms * 1e-3) / 1e12, 'peak': tf_burst, 'unit': 'TFLOP/s',
            'frac': tfl / (tms * 1e-3) / 1e12 / tf_burst, 'ms_per_launch': tms,
            'traffic': tj.get('tsa_gemm', {}).get('dram_bytes_per_launch') if (B == 256 and N == 2048 and same_rev) else None,
            'ncu_tensor_pipe_active_pct': tj.get('tsa_gemm', {}).get('tensor_pipe_active_pct')}
        # the chained form of the same pooling step (gp_pool_chain_bf16: A' = S^T A S in one launch, T on chip) next to
        # the two launches the step uses by default (T = S^T A with CTA pairs, A' = T S); DESIGN.md section 4 explains why
        # the chain is the slower one at K = 512 (TMEM leaves it 128-column T tiles on single CTAs: shared-memory port)
        if K0 <= 512 and K0 % 8 == 0:
            import ctypes as C_
            from graph_pooling_b200._lib import call as call_
            apf = torch.empty(B, K0, K0, device=dev)
            apb = T.bfbuf(wsb, B, K0, K0)

            def chain(keep_t):
                call_('gp_pool_chain_bf16', sop.ptr, C_.c_longlong(sop.ld), adjb.ptr, C_.c_longlong(adjb.ld),
                      nbd.data_ptr(), None, B, N, K0, tbuf.ptr if keep_t else None,
                      C_.c_longlong(tbuf.ld if keep_t else 0), apf.data_ptr(), C_.c_longlong(K0), apb.ptr,
                      C_.c_longlong(apb.ld), None)

            def two():
                tsa()
                T.tcgemm(tbuf, T.KM, sop, T.MN, K0, K0, N, B, Cf=(apf.data_ptr(), K0, K0 * K0), Cb=apb,
                         lim=nbd.data_ptr(), lim_k=1)
            for f_ in (lambda: chain(True), lambda: chain(False), two):
                for _ in range(2):
                    f_()
            cfl = tfl + float(np.sum(2.0 * K0 * K0 * nbf))
            c_t, c_n, c_2 = [timed_local(f_, reps) / reps for f_ in (lambda: chain(True), lambda: chain(False), two)]
            roof['chained_pooling'] = {
                'kernel': "gp::chain::pool_chain_kernel<%d> (A' = S^T A S, T = S^T A in TMEM -> bf16 shared tile -> second "
                          "tcgen05.mma; %s)" % (2 if K0 > 256 else 1, 'two-CTA cluster, DSMEM exchange of T tiles'
                                                if K0 > 256 else 'one CTA per row block'),
                'flops': cfl, 'ms_chained_T_stored_once': c_t, 'ms_chained_T_on_chip_only': c_n,
                'ms_two_launches_default': c_2, 'tflops_chained': cfl / (c_t * 1e-3) / 1e12,
                'tflops_two_launches': cfl / (c_2 * 1e-3) / 1e12, 'default': 'two launches (GP_CHAIN=1 selects the chain)'}
            del apf, apb
        del sprob, tbuf

    # ---- the step against ITS roof (top-level `roofline`); the kernel timed alone above is `dominant_kernel` ------
    if prec == 'bf16':        # model AI >> ridge: the step is tensor-bound; sustained peak (kernels inside a long step)
        step_roof = {'bound': 'tensor', 'achieved': roof['step_tflops'], 'peak': tf_sus, 'unit': 'TFLOP/s',
                     'frac': roof['step_tflops'] / tf_sus,
                     'what': 'whole train step: algorithmic TFLOP (roofline.py: real n_b blocks, reference association, '
                             'fwd+bwd) / device time, against the SUSTAINED bf16 peak'}
    else:                     # ENZYMES / DD-base sized work: HBM-bound (AI ~ 35 flop/B); bytes counted at fp32
        step_roof = {'bound': 'hbm', 'achieved': roof['step_gbs'], 'peak': hbm, 'unit': 'GB/s',
                     'frac': roof['step_gbs'] / hbm,
                     'what': 'whole train step: compulsory bytes (roofline.py, fp32: adjacency read fwd + bwd, saved '
                             'activations written once and read once) / device time'}
    step_roof['traffic'] = roof.get('traffic')
    step_roof['traffic_note'] = ('dram__bytes_read + dram__bytes_write of ONE launch of the dominant kernel '
                                 '(dominant_kernel: its algorithmic bytes are algorithmic_bytes_per_launch there), from '
                                 'ncu --set full on the same kernel source (profiles/r2_traffic.json); not a step total')
    step_roof['peak_source'] = src + ' (MEASURED_PEAKS.json)'
    step_roof['kernels'] = kernels_tab
    step_roof['kernels_note'] = ('one extra untimed step, every C-ABI call bracketed by CUDA events (profile.py); rows = '
                                 '(entry point, shape) groups with >= 3 %% of the %.3f ms of calls; flops / bytes are '
                                 'algorithmic (dense extents) at the dtype each call reads and writes'
                                 % (prof_ms if prof_ms else 0.0))
    step_roof['dominant_kernel'] = roof

    # ---- comparators ------------------------------------------------------------------------------
    cpu = None
    aten = None
    if not args.no_cpu_baseline and world == 1:       # rank 0 at N = 1 only (other ranks would contend for the cores)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = args.cpu_sample or default_cpu_sample(cfg)
        v, cms, desc = time_oracle(args.workload, sample, 2, 1, args.seed, 'cpu', True)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc, 'ms_per_step': cms}
        v2, cms2, desc2 = time_oracle(args.workload, sample, 2, 1, args.seed, 'cpu', False)
        cpu['fwd_bwd_only'] = {'value': v2, 'ms_per_step': cms2, 'sample': desc2}
        try:                  # the same oracle in torch eager ON THE B200 (cuBLAS / ATen kernels): "ATen-on-B200"
            del xin, u
            torch.cuda.empty_cache()
            gs_ = int(min(B, 32 if cfg['N'] >= 1024 else 4096))
            v3, ams, desc3 = time_oracle(args.workload, gs_, 3, 2, args.seed, dev, True)
            aten = {'value': v3, 'unit': UNIT, 'ms_per_step': ams, 'sample': desc3}
        except Exception as ex:
            aten = {'error': str(ex)[:200]}
    enz = None
    if world == 1 and args.workload == 'cfg4_diffpool_256x2048' and not os.environ.get('GP_BENCH_NO_ENZ'):
        try:
            enz = enzymes_regime_point(dev, hbm)
        except Exception as ex:
            enz = {'error': str(ex)[:200]}

    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': prec,
           'data': 'real: bundled ENZYMES graphs (tests/golden/dataset_enzymes.npz, made through the reference loader)'
           if cfg.get('fixture') else 'synthetic',
           'config': {'workload': args.workload, 'precision': prec, 'graphs_per_gpu_per_step': B, 'nodes': cfg['N'],
                      'hidden': cfg['H'], 'assign_ratio': cfg['ratio'], 'num_pooling': cfg['P'],
                      'step': 'zero_grad+forward+loss(CE+linkpred)+backward+clip_grad_norm+Adam (train.py:196-210)',
                      'l2': 'inputs larger than L2 (adjacency %.1f GB)' % (adj.numel() * 4 / 1e9)
                      if adj.numel() * 4 > 126e6 else 'inputs smaller than L2; not flushed',
                      'cuda_graph': bool(use_graph),
                      'parallelism': 'dp%d' % world},
           'clocks': clocks, 'e2e': e2e, 'e2e_u8_feed': e2e_u8, 'e2e_edge_feed': e2e_edges, 'gpu_launches': launches, 'roofline': step_roof,
           'cpu_baseline': cpu, 'aten_on_b200': aten, 'enzymes_regime': enz, 'dp_check': dp_check}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
