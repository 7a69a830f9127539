"""Worker of tests/test_gpu_dp.py::test_nccl_two_ranks (launched by torch.distributed.run, one process per GPU).

Checks on the CUDA / NCCL path of dp.py:
  1. the all-reduced flat gradient equals the mean of the shard gradients (gathered from every rank), and equals the
     mean of the ORACLE's fp64 gradients on the same shards (each GPU == the reference on its shard);
  2. after several FlatAdam steps the replicas are bit-identical (MAX - MIN of every parameter over ranks == 0).
Prints one JSON line on rank 0; exit code != 0 on failure."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from graph_pooling_b200 import dp, encoders
    from helpers import synth_batch
    from oracle import diffpool_oracle as orc
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    precision = int(os.environ.get('GP_DP_PRECISION', '0'))
    N, D, H, C, B = (48, 5, 16, 3, 4 * world) if precision == 0 else (256, 16, 32, 2, 2 * world)
    torch.manual_seed(3)
    mo = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
    mc = encoders.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
    mc.load_state_dict(mo.state_dict())
    mc = mc.to(dev)
    mc.precision = precision
    x, adj, nb, label = synth_batch(9, B, N, D, 3, N, C, 0.12)
    sh = dp.shard_batch(rank, world, torch.tensor(x), torch.tensor(adj), nb, torch.tensor(label))
    xc, ac, lc = sh['x'].to(dev), sh['adj'].to(dev), sh['label'].to(dev)

    opt = dp.FlatAdam(list(mc.parameters()), lr=1e-3, clip=2.0)
    flat = opt.grads.attach(mc)
    flat.zero()
    yp = mc(xc, ac, sh['nb'], assign_x=xc)
    loss = mc.loss(yp, lc, ac, sh['nb'])
    loss.backward()
    local_grad = flat.flat.clone()
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    mean = torch.stack(gathered).double().mean(0)
    flat.all_reduce(average=True)
    flat.apply_pending_scale()
    reduced = flat.flat.double()
    err_mean = float((reduced - mean).norm() / mean.norm())

    # the oracle on every shard (fp64, CPU), gradients averaged
    acc = None
    for r in range(world):
        s = dp.shard_batch(r, world, torch.tensor(x), torch.tensor(adj), nb, torch.tensor(label))
        m = orc.SoftPoolingGcnEncoder(N, D, H, H, C, 3, H, assign_ratio=0.25)
        m.load_state_dict(mo.state_dict())
        m = m.double()
        orc.train_step(m, s['x'].double(), s['adj'].double(), s['label'], s['nb'])
        g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        acc = g if acc is None else acc + g
    want = (acc / world)
    err_oracle = float((reduced.cpu() - want).norm() / want.norm())

    # a few optimiser steps, then replicas must be bit-identical
    for _ in range(3):
        flat.zero()
        yp = mc(xc, ac, sh['nb'], assign_x=xc)
        loss = mc.loss(yp, lc, ac, sh['nb'])
        loss.backward()
        flat.all_reduce(average=True)
        opt.step()
    pmax, pmin = opt.flat_p.clone(), opt.flat_p.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    spread = float((pmax - pmin).abs().max())
    torch.cuda.synchronize()
    ok = err_mean < 1e-6 and spread == 0.0 and err_oracle < (2e-4 if precision == 0 else 0.1)
    if rank == 0:
        print(json.dumps({'world': world, 'precision': precision, 'reduced_vs_mean_of_shards': err_mean,
                          'reduced_vs_oracle_fp64': err_oracle, 'replica_param_spread': spread, 'ok': ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
